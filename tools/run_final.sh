#!/bin/bash
# final verification: full GPU suite, smoke(), default bench line, reference arm
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['single_stream'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline'])
print(d['train'])
r=json.loads(open('gpurun_out/bench_ref.log').read().strip().splitlines()[-1])
print({k:r[k] for k in ('impl','value','ms_per_step','cpu_baseline','e2e')})
PY
