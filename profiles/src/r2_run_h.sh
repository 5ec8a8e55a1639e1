#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_rollout_gpu.py -q -x -k "fused or bf16_backward or golden" > gpurun_out/h_pytest.txt 2>&1; echo "tests1 exit $?" > gpurun_out/h.log
timeout 120 python profiles/src/r2_quick.py > gpurun_out/h_quick.txt 2>&1
tail -3 gpurun_out/h_pytest.txt; cat gpurun_out/h.log gpurun_out/h_quick.txt
