#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --workload train --batch 8 --steps 2 --warmup 3 > gpurun_out/train_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches_b8.csv python bench.py --workload train --batch 8 --steps 2 --warmup 3 > gpurun_out/train_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/train_plain.log | cut -c1-400
