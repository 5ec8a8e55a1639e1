#!/bin/bash
for v in "$@"; do
  echo "== $v"; RSSM_ROLLOUT_LIB=scratch/lib_$v.so python scratch/exp_fwd.py 37888 2>&1 | grep -v "^$" | tr '\n' ';'; echo
done
