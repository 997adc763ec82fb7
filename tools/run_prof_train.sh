#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B=${1:-8}
timeout 600 python bench.py --workload train --batch $B --steps 2 --warmup 3 > gpurun_out/train_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches_b$B.csv python bench.py --workload train --batch $B --steps 2 --warmup 3 > gpurun_out/train_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/train_plain.log | cut -c1-300
