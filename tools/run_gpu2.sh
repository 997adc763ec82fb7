#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --workload train --steps 10 --warmup 3 > gpurun_out/g2_train.log 2> gpurun_out/g2_train.err; echo "train rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload adversarial --steps 6 --warmup 3 > gpurun_out/g2_adv.log 2> gpurun_out/g2_adv.err; echo "adv rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload deeplab --steps 5 --warmup 3 > gpurun_out/g2_dl.log 2> gpurun_out/g2_dl.err; echo "deeplab rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 10 --no-train > gpurun_out/g2_inf.log 2> gpurun_out/g2_inf.err; echo "infer rc=$?"
for f in g2_train g2_adv g2_dl g2_inf; do tail -1 gpurun_out/$f.log | cut -c1-330; tail -2 gpurun_out/$f.err | cut -c1-300; done
