#!/bin/bash
# usage: run_scale.sh N — the driver's launch for N GPUs of one box: default bench line (DP-training headline under --gpus N>1)
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; echo "rc=$?"
tail -1 gpurun_out/r02_bench_n${N}.json | cut -c1-600
tail -3 gpurun_out/r02_bench_n${N}.err | cut -c1-300
