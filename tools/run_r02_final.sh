#!/bin/bash
# Round-2 final evidence on ONE B200: default bench line + reference arm, ncu launch lists of one inference frame and one
# training step, and one `ncu --set full` capture of the persistent conv kernel on the training step (each ncu pass only
# after the same command exited 0 without it).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log
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
CMD2="python bench.py --steps 3 --warmup 3 --no-train --no-extra --no-comparator --lanes 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 88 -c 22 -f -o gpurun_out/r02_conv_tc_infer $CMD2 > gpurun_out/r02_convfull_ncu.log 2>&1
echo "conv full rc=$?"
ncu -i gpurun_out/r02_conv_tc_infer.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,launch__cluster_size,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active > gpurun_out/r02_conv_tc_infer_ncu_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_conv_tc_infer.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tcp -s 60 -c 12 -f -o gpurun_out/r02_conv_tcp_train $CMD3 > gpurun_out/r02_tcpfull_ncu.log 2>&1
echo "tcp full rc=$?"
ncu -i gpurun_out/r02_conv_tcp_train.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,launch__registers_per_thread,launch__block_size,launch__grid_size,smsp__inst_executed.sum > gpurun_out/r02_conv_tcp_train_ncu_full.csv 2>/dev/null
rm -f gpurun_out/r02_conv_tcp_train.ncu-rep
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline'])
print(d.get('train'))
PY
