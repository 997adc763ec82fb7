#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/t_all.log
timeout 300 python bench.py --workload train --batch 8 --steps 60 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --workload deeplab --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deeplab', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --steps 100 --warmup 10 --no-train --no-extra --no-comparator --lanes 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('infer', d['value'], d['ms_per_step'], d['device_timed']['ms_per_step'])"
grep deterministic:bisenet -A12 gpurun_out/r02_parity.json | head -20
