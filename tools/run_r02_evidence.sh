#!/bin/bash
# Round-2 evidence on ONE B200 (gpurun --timeout 2700 -- 'bash tools/run_r02_evidence.sh'): the full -m gpu suite (writes
# gpurun_out/r02_parity.json), smoke(), the default bench line + reference arm, then the ncu launch lists of one inference
# run and one training run and the --set full capture of the conv launches (each only after the plain command exited 0).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1300 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference_n1.json 2> /dev/null; echo "reference rc=$?"
CMD="python bench.py --steps 20 --warmup 5 --no-train --no-extra --no-comparator --lanes 1"
timeout 300 $CMD > gpurun_out/r02_infer_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_infer_launches.csv $CMD > gpurun_out/r02_infer_ncu.log 2>&1
echo "launch list rc=$?"
CMD3="python bench.py --workload train --batch 8 --steps 2 --warmup 3"
timeout 300 $CMD3 > gpurun_out/r02_train_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_train_launches.csv $CMD3 > gpurun_out/r02_train_ncu.log 2>&1
echo "train launch list rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('single_stream'), d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline'])
print(d.get('train'))
PY
