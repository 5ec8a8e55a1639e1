#!/bin/bash
cd "$GRAFT_REPO_ROOT"
RSSM_ROLLOUT_LIB=profiles/src/lib_timing.so RSSM_FZ_TIMING=1 python profiles/src/r2_fz_timing.py > gpurun_out/b_timing.txt 2>&1
tail -120 gpurun_out/b_timing.txt
