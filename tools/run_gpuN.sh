#!/bin/bash
# usage: run_gpuN.sh N  -- the driver's scaling protocol (default bench line = inference + train sub-object) plus the train workload
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N > gpurun_out/g${N}_default.log 2> gpurun_out/g${N}_default.err; echo "default rc=$?"
timeout 600 $TR bench.py --gpus $N --workload train --steps 20 --warmup 5 > gpurun_out/g${N}_train.log 2> gpurun_out/g${N}_train.err; echo "train rc=$?"
timeout 600 $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/g${N}_ref.log 2> gpurun_out/g${N}_ref.err; echo "ref rc=$?"
for f in g${N}_default g${N}_train g${N}_ref; do tail -1 gpurun_out/$f.log | cut -c1-400; tail -2 gpurun_out/$f.err | cut -c1-300; done
