for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_bisenet.py -x -q -k "test_train_forward_vs_reference_golden" 2>&1 | grep -E "assert|passed|failed|Error" | head -5; done
echo NOCLUSTER
for i in 1 2 3; do RTSDS_NO_CLUSTER_SPLITK=1 timeout 300 python -m pytest tests/test_gpu_bisenet.py -x -q -k "test_train_forward_vs_reference_golden" 2>&1 | grep -E "assert|passed|failed|Error" | head -5; done
