#!/bin/bash
# Round-2 profiling visit (GPU box, from the repo root): phase split, per-shape kernel tables, the ncu launch list of one
# bench.py step (BAIR headline) and ncu --set full captures of the top kernels.  Each ncu pass runs only after the same
# command has exited 0 without ncu.  Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
python tools/time_phases.py bair 32 2>&1 | grep -E "^(condition|ddim|decode|full)" > gpurun_out/phases_bair_r2.log; cat gpurun_out/phases_bair_r2.log
python tools/kernel_table.py 32 bair > gpurun_out/ktable_bair_r2.md 2> gpurun_out/ktable_bair_r2.err; tail -1 gpurun_out/ktable_bair_r2.md
python tools/kernel_table.py 32 smmnist > gpurun_out/ktable_smmnist_r2.md 2> gpurun_out/ktable_smmnist_r2.err; tail -1 gpurun_out/ktable_smmnist_r2.md
ARGS="--steps 1 --warmup 3 --no-cpu-baseline --no-per-config --no-gpu-reference --ncu-range"
python bench.py $ARGS > gpurun_out/plain_list.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 12000 --csv \
    --log-file gpurun_out/launches_r2_bair.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
tail -n 2 gpurun_out/ncu_list.log | cut -c 1-300; wc -l gpurun_out/launches_r2_bair.csv
KEYS=("stw_fused C=64 12x32x32" "temporal_fused" "gemm rows=81920 n=64 k=6400 taps=25" "gemm rows=393216 n=64 k=576" "groupnorm_apply C=64" "gemm rows=327680 n=64 k=832 taps=13" "gemm rows=10240 n=64 k=832 taps=13" "init_corner_fix" "window_attention hid=256 12x16x16")
python tools/ncu_target.py --dataset bair "${KEYS[@]}" > gpurun_out/ncu_plain_r2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r2_bair \
    python tools/ncu_target.py --dataset bair "${KEYS[@]}" > gpurun_out/ncu_r2.log 2>&1
tail -n 3 gpurun_out/ncu_plain_r2.log; tail -n 3 gpurun_out/ncu_r2.log; ls -la gpurun_out/*.ncu-rep
