timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bisenet.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-train 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['single_stream']['value'], d['single_stream']['ms_per_step'], d['e2e']['value'], d['e2e']['serial_fps'])"
