#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --steps 20 --warmup 5 --no-train > gpurun_out/infer_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/infer_launches.csv python bench.py --steps 20 --warmup 5 --no-train > gpurun_out/infer_ncu.log 2>&1; echo "ncu rc=$?"
tail -1 gpurun_out/infer_plain.log | cut -c1-400
