timeout 900 python -m pytest tests/test_gpu_conv.py -x -q -k many_tiles 2>&1 | tail -2
timeout 300 python bench.py --workload train --batch 8 --steps 10 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stream', d['value'], d['ms_per_step'], d['final_loss'])"
timeout 300 python bench.py --workload deeplab 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('deeplab stream', d['value'], d['ms_per_step'])"
RTSDS_NO_HALO_STREAM=1 timeout 300 python bench.py --workload deeplab 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('deeplab nostream', d['value'], d['ms_per_step'])"
