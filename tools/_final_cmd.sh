python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1.log 2>&1; tail -1 gpurun_out/smoke_r1.log
tools/gpu_round.sh r1w "stw_fused C=64 30x32x32" "step gemm rows=983040 n=64 k=576" "step gemm rows=655360 n=64 k=12544" "step gemm rows=983040 n=64 k=1152" "groupnorm_apply C=64 P=30720" "temporal_fused" "decode gemm rows=245760 n=256 k=2304" "stw_fused C=128 30x16x16"
