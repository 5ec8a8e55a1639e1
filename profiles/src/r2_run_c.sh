#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 400 python -m pytest tests/test_rollout_gpu.py tests/test_models_gpu.py tests/test_bench_configs_gpu.py -q -x > gpurun_out/c_pytest.txt 2>&1; echo "tests exit $?" > gpurun_out/c.log
timeout 120 python profiles/src/r2_quick.py > gpurun_out/c_quick.txt 2>&1
timeout 120 python profiles/src/r2_quick.py --no-prior >> gpurun_out/c_quick.txt 2>&1
RSSM_FWD_ONE_WARP=1 timeout 120 python profiles/src/r2_quick.py >> gpurun_out/c_quick.txt 2>&1
tail -5 gpurun_out/c_pytest.txt; cat gpurun_out/c.log gpurun_out/c_quick.txt
