K='step gemm rows=983040 n=64 k=576'
python tools/ncu_target.py "$K" > gpurun_out/ncu_plain_h.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_halo2 python tools/ncu_target.py "$K" > gpurun_out/ncu_h.log 2>&1
tail -n 3 gpurun_out/ncu_h.log
