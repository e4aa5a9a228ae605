#!/bin/bash
# Run-to-run spread of one configuration: N fresh processes, SM clock / power sampled every 100 ms beside each.
NAME=${1:-smmnist}; N=${2:-6}
for i in $(seq 1 $N); do
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 100 > /tmp/smi_$i.log &
  SMI=$!
  python tools/bench_configs.py $NAME 2>&1 | grep "$NAME"
  kill $SMI
  sort /tmp/smi_$i.log | uniq -c | sort -rn | head -4
done
