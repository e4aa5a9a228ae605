# ncu --set full of the top kernels (the recipe behind profiles/ncu_full_r2.md); run on the GPU box from the repo root
KEYS=("stw_fused C=64 30x32x32" "gemm rows=983040 n=64 k=576" "gemm rows=655360 n=64 k=12544" "gemm rows=983040 n=64 k=1152" "groupnorm_apply C=64 P=30720" "temporal_fused")
python tools/ncu_target.py "${KEYS[@]}" > gpurun_out/ncu_plain_r2.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r2 \
    python tools/ncu_target.py "${KEYS[@]}" > gpurun_out/ncu_r2.log 2>&1
tail -n 3 gpurun_out/ncu_plain_r2.log; tail -n 3 gpurun_out/ncu_r2.log
