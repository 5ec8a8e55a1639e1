#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 400 python -m pytest tests/test_rollout_gpu.py tests/test_bench_configs_gpu.py -q -x > gpurun_out/d_pytest.txt 2>&1; echo "tests exit $?" > gpurun_out/d.log
timeout 120 python profiles/src/r2_quick.py > gpurun_out/d_quick.txt 2>&1
RSSM_ROLLOUT_LIB=profiles/src/lib_timing.so RSSM_FZ_TIMING=1 timeout 120 python profiles/src/r2_fz_timing.py > gpurun_out/d_timing.txt 2>&1
tail -4 gpurun_out/d_pytest.txt; cat gpurun_out/d.log gpurun_out/d_quick.txt; grep -A14 "fz timing B=256\|fz timing B=37888" gpurun_out/d_timing.txt | tail -64
