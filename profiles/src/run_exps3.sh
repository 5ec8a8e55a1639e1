#!/bin/bash
for v in "$@"; do
  echo "== $v"; RSSM_ROLLOUT_LIB=profiles/src/lib_$v.so python profiles/src/exp_fwd.py 37888 2>&1 | grep -v "^$" | tr '\n' ';'; echo
done
