#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for i in 1 2; do
echo "== tiled record (default), pass $i" >> gpurun_out/n_quick.txt
timeout 120 python profiles/src/r2_quick.py >> gpurun_out/n_quick.txt 2>&1
echo "== row-layout record (RSSM_REC_ROW_LAYOUT=1), pass $i" >> gpurun_out/n_quick.txt
RSSM_REC_ROW_LAYOUT=1 timeout 120 python profiles/src/r2_quick.py >> gpurun_out/n_quick.txt 2>&1
done
cat gpurun_out/n_quick.txt
