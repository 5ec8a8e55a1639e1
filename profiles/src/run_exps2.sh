#!/bin/bash
for v in noprefetch np_r3 np_r4; do
  for B in 16384 28416 37888; do
    echo "== $v B=$B"; RSSM_ROLLOUT_LIB=profiles/src/lib_$v.so python profiles/src/exp_fwd.py $B 2>&1 | grep -v "^$" | tr '\n' ';'; echo
  done
done
