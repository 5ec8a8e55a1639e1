#!/bin/bash
# same-box timing of several builds of the library (profiles/src/lib_NAME.so, RSSM_ROLLOUT_LIB); "main" = the shipped library
# usage: r2_run_variants.sh TAG NAME...     (file gpurun_out/TAG_variants.txt)
cd "$GRAFT_REPO_ROOT"
P=gpurun_out/${1:-var}; shift
: > ${P}_variants.txt
for round in 1 2; do
for v in "$@"; do
  echo "== $v (round $round)" >> ${P}_variants.txt
  if [ "$v" = main ]; then timeout 200 python profiles/src/r2_quick.py --ab-sizes >> ${P}_variants.txt 2>&1
  else RSSM_ROLLOUT_LIB=profiles/src/lib_$v.so timeout 200 python profiles/src/r2_quick.py --ab-sizes >> ${P}_variants.txt 2>&1; fi
done; done
cat ${P}_variants.txt
