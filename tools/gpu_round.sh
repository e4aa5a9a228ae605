#!/bin/bash
# One GPU-box visit: parity tests, bench, per-shape kernel table and (optionally) ncu --set full captures.
# usage: tools/gpu_round.sh TAG [ncu keys...]
TAG=$1; shift
python -m pytest tests -m gpu -q > gpurun_out/tests_$TAG.log 2>&1; tail -15 gpurun_out/tests_$TAG.log
python bench.py --steps 2 --warmup 3 --profile-kernels > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_$TAG.err; cut -c 1-400 gpurun_out/bench_$TAG.json
python tools/kernel_table.py > gpurun_out/ktable_$TAG.md 2> gpurun_out/ktable_$TAG.err; tail -3 gpurun_out/ktable_$TAG.err
if [ $# -gt 0 ]; then
  python tools/ncu_target.py "$@" > gpurun_out/ncu_plain_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_$TAG \
      python tools/ncu_target.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
  tail -5 gpurun_out/ncu_plain_$TAG.log; tail -5 gpurun_out/ncu_$TAG.log
fi
