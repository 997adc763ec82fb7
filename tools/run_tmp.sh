#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 0 1 0 1; do
if [ $v = 1 ]; then export RTSDS_NO_BN_FUSED_APPLY=1; else unset RTSDS_NO_BN_FUSED_APPLY; fi
timeout 300 python bench.py --workload train --batch 8 --steps 60 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train nofuse=$v', d['value'], d['ms_per_step'], d.get('launches_per_step'))"
done
for v in 0 1; do
if [ $v = 1 ]; then export RTSDS_NO_BN_FUSED_APPLY=1; else unset RTSDS_NO_BN_FUSED_APPLY; fi
timeout 300 python bench.py --workload deeplab --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deeplab nofuse=$v', d['value'], d['ms_per_step'], d.get('launches_per_step'))"
done
