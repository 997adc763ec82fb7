#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_all.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "bench1 rc=$?"
tail -c 600 gpurun_out/bench1.err
python -c "
import json; d=json.loads(open('gpurun_out/bench1.log').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']); print(d['train'])"
