#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun --timeout 2400 -- 'bash tools/run_r02.sh'): default bench line + reference arm,
# then the ncu launch list of one inference forward and the --set full capture of its conv launches.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_n1.json 2> /dev/null; echo "reference rc=$?"
CMD="python bench.py --steps 20 --warmup 5 --no-train --no-extra --no-comparator --lanes 1"
$CMD > gpurun_out/r02_infer_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_infer_launches.csv $CMD > gpurun_out/r02_infer_ncu.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 3 --warmup 3 --no-train --no-extra --no-comparator --lanes 1"
$CMD2 > gpurun_out/r02_convfull_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 24 -f -o gpurun_out/r02_conv_tc_infer $CMD2 > gpurun_out/r02_convfull_ncu.log 2>&1
echo "conv full rc=$?"
CMD3="python bench.py --workload train --batch 8 --steps 2 --warmup 3"
$CMD3 > gpurun_out/r02_train_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_train_launches.csv $CMD3 > gpurun_out/r02_train_ncu.log 2>&1
echo "train launch list rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/hbm_kernels.py > gpurun_out/r02_hbm_kernels.json 2> gpurun_out/r02_hbm_kernels.err &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:_kernel --csv --log-file gpurun_out/r02_hbm_ncu.csv python tools/hbm_kernels.py --ncu > gpurun_out/r02_hbm_ncu.log 2>&1
echo "hbm table rc=$?"
