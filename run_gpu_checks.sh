#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q --maxfail=40 -m gpu > gpurun_out/t_all.log 2>&1; echo "all rc=$?"
grep -E "^E +Assert|passed|failed|^FAILED|Error|rtsds:" gpurun_out/t_all.log | cut -c1-300 | head -30
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --batch 4 > gpurun_out/train_b4.log 2> gpurun_out/train_b4.err; echo "train rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/train_b4.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['launches_per_step'], d['roofline']['frac'])"
tail -3 gpurun_out/train_b4.err
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['launches_per_step'], d['readme_protocol'], d['latency_cold_l2_ms'])"
