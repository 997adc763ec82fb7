#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q --maxfail=40 -m gpu > gpurun_out/t_all.log 2>&1; echo "all rc=$?"
grep -E "^E +Assert|passed|failed|^FAILED|Error" gpurun_out/t_all.log | cut -c1-300 | head -40
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
