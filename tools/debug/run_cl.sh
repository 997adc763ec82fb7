timeout 600 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_bisenet.py tests/test_gpu_deeplab.py tests/test_gpu_disc.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-train 2>&1 | tail -1 | cut -c1-200
RTSDS_NO_CLUSTER_SPLITK=1 timeout 300 python bench.py --no-train 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --workload train --batch 8 --steps 10 --warmup 5 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --workload deeplab 2>&1 | tail -1 | cut -c1-200
