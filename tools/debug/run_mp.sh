timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k maxpool 2>&1 | tail -3
python bench.py --workload train --batch 8 --steps 10 --warmup 5 2>&1 | tail -1 | cut -c1-250
