#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 120 python profiles/src/r2_packed_exp.py > gpurun_out/i_packed.txt 2>&1
RSSM_EXP_ROW_PITCH=368 timeout 120 python profiles/src/r2_packed_exp.py >> gpurun_out/i_packed.txt 2>&1
RSSM_EXP_ROW_PITCH=384 timeout 120 python profiles/src/r2_packed_exp.py >> gpurun_out/i_packed.txt 2>&1
cat gpurun_out/i_packed.txt
