#!/bin/bash
for v in base nostores noprefetch nomufu; do
  for B in 16384 32768; do
    echo "== $v B=$B"; RSSM_ROLLOUT_LIB=profiles/src/lib_$v.so python profiles/src/exp_fwd.py $B 2>&1 | grep -v "^$" | tr '\n' ';'; echo
  done
done
