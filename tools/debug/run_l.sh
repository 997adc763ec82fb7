timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bisenet.py tests/test_gpu_deeplab.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload train --batch 8 --steps 10 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['final_loss'], d['e2e']['value'])"
