#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_rollout_gpu.py tests/test_bench_configs_gpu.py tests/test_models_gpu.py -q -x > gpurun_out/m_pytest.txt 2>&1; echo "tests exit $?" > gpurun_out/m.log
timeout 120 python profiles/src/r2_quick.py > gpurun_out/m_quick.txt 2>&1
tail -3 gpurun_out/m_pytest.txt; cat gpurun_out/m.log gpurun_out/m_quick.txt
