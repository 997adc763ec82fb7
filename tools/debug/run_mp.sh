timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bisenet.py tests/test_gpu_deeplab.py -x -q 2>&1 | tail -3
python bench.py --workload train --batch 8 --steps 10 --warmup 5 2>&1 | tail -1 | cut -c1-250
