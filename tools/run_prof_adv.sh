#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --workload adversarial --steps 2 --warmup 3 > gpurun_out/adv_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/adv_launches.csv python bench.py --workload adversarial --steps 2 --warmup 3 > gpurun_out/adv_ncu.log 2>&1; echo "ncu rc=$?"
tail -1 gpurun_out/adv_plain.log | cut -c1-300
