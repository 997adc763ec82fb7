#!/bin/bash
# First-contact GPU run: safe kernels first, then the tcgen05 conv, then end to end.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --maxfail=40 -m gpu > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 300 python tests/debug_tc.py > gpurun_out/debug_tc.log 2>&1; echo "debug_tc rc=$?"
timeout 600 python -m pytest tests/test_gpu_conv.py -q --maxfail=60 -m gpu > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?"
timeout 600 python -m pytest tests/test_gpu_bisenet.py -q --maxfail=20 -m gpu > gpurun_out/t_bisenet.log 2>&1; echo "bisenet rc=$?"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/t_kernels.log; tail -30 gpurun_out/debug_tc.log; tail -5 gpurun_out/t_conv.log; tail -5 gpurun_out/t_bisenet.log; tail -3 gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
