python -m pytest tests/test_gpu_bn_bwd.py -x -q 2>&1 | tail -5
python tools/bn_micro.py 2>&1 | head -4
python bench.py --workload train --batch 8 --steps 10 --warmup 5 2>&1 | tail -1 | cut -c1-250
