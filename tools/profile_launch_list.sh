# ncu launch list of one bench.py step (the recipe behind profiles/launches_r2.md); run on the GPU box from the repo root
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ncu-range > gpurun_out/plain_list.log 2>&1 && \
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 9000 --csv \
    --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ncu-range > gpurun_out/ncu_list.log 2>&1
tail -n 2 gpurun_out/ncu_list.log | cut -c 1-300; wc -l gpurun_out/launches_r2.csv
