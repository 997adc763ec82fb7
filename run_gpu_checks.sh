#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 rc=$?"
tail -3 gpurun_out/bench2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench2.log').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']); print(d['train'])"
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench1.log 2> gpurun_out/bench1.err; echo "bench1 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench1.log').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']); print(d['train'])"
