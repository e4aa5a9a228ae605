KEYS=("stw_fused C=64 12x32x32" "temporal_fused" "gemm rows=81920 n=64 k=6400 taps=25" "gemm rows=393216 n=64 k=576" "groupnorm_apply C=64" "gemm rows=327680 n=64 k=832 taps=13" "gemm rows=10240 n=64 k=832 taps=13" "init_corner_fix" "window_attention hid=256 12x16x16")
python tools/ncu_target.py --dataset bair "${KEYS[@]}" > gpurun_out/ncu_plain_r2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r2_bair \
    python tools/ncu_target.py --dataset bair "${KEYS[@]}" > gpurun_out/ncu_r2.log 2>&1
tail -n 9 gpurun_out/ncu_plain_r2.log
