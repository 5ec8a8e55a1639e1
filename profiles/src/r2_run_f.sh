#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_rollout_gpu.py -q -x -k "fused or bf16_backward or golden" > gpurun_out/f_pytest.txt 2>&1; echo "tests1 exit $?" > gpurun_out/f.log
timeout 300 python -m pytest tests/test_bench_configs_gpu.py tests/test_models_gpu.py -q -x > gpurun_out/f_pytest2.txt 2>&1; echo "tests2 exit $?" >> gpurun_out/f.log
timeout 120 python profiles/src/r2_quick.py > gpurun_out/f_quick.txt 2>&1
RSSM_BWD_TWO_WARP=1 timeout 120 python profiles/src/r2_quick.py >> gpurun_out/f_quick.txt 2>&1
RSSM_ROLLOUT_LIB=profiles/src/lib_timing.so RSSM_FZ_TIMING=1 timeout 120 python profiles/src/r2_fz_timing.py > gpurun_out/f_timing.txt 2>&1
tail -4 gpurun_out/f_pytest.txt; tail -4 gpurun_out/f_pytest2.txt; cat gpurun_out/f.log gpurun_out/f_quick.txt; grep -A12 "fz timing B=256\|fz timing B=37888" gpurun_out/f_timing.txt | tail -90
