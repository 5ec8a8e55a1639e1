#!/bin/bash
# same-box timing of library builds with the two kernels alternating (pipeline mode of r2_quick.py)
cd "$GRAFT_REPO_ROOT"
P=gpurun_out/${1:-pipe}; shift
: > ${P}.txt
for round in 1 2; do
for v in "$@"; do
  echo "== $v (round $round)" >> ${P}.txt
  if [ "$v" = main ]; then timeout 200 python profiles/src/r2_quick.py --only-bench --pipeline >> ${P}.txt 2>&1
  else RSSM_ROLLOUT_LIB=profiles/src/lib_$v.so timeout 200 python profiles/src/r2_quick.py --only-bench --pipeline >> ${P}.txt 2>&1; fi
done; done
cat ${P}.txt
