K='stw_fused C=64 30x32x32'
EXTDM_STW_TC=1 python tools/ncu_target.py "$K" > gpurun_out/ncu_plain_t.log 2>&1 && \
EXTDM_STW_TC=1 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_stwtc python tools/ncu_target.py "$K" > gpurun_out/ncu_t.log 2>&1
tail -n 3 gpurun_out/ncu_t.log
