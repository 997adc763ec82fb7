#!/bin/bash
# usage: run_ncu_full.sh <out-name> <kernel-regex> <count> -- <cmd...>
mkdir -p gpurun_out
name=$1; regex=$2; cnt=$3; shift 4
"$@" > gpurun_out/${name}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$regex -c $cnt -o gpurun_out/$name -f "$@" > gpurun_out/${name}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${name}_ncu.log; cat gpurun_out/${name}_plain.log | tail -12
