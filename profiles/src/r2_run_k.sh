#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_rollout_gpu.py -q -x -k "fused or bf16_backward or golden or projected" > gpurun_out/k_pytest.txt 2>&1; echo "tests1 exit $?" > gpurun_out/k.log
timeout 120 python profiles/src/r2_quick.py > gpurun_out/k_quick.txt 2>&1
tail -3 gpurun_out/k_pytest.txt; cat gpurun_out/k.log gpurun_out/k_quick.txt
