#!/bin/bash
# same-box A/B of two builds of the library: GPU tests of the current build, then kernel timings of both
# usage: r2_run_ab.sh TAG [BASE_LIB]   (files gpurun_out/TAG_*)
cd "$GRAFT_REPO_ROOT"
P=gpurun_out/${1:-ab}
BASE=${2:-profiles/src/lib_base.so}
timeout 900 python -m pytest tests -m gpu -q -x > ${P}_pytest.txt 2>&1; echo "tests exit $?" > ${P}.log
tail -3 ${P}_pytest.txt
timeout 300 python profiles/src/r2_quick.py > ${P}_new.txt 2>&1; echo "new exit $?" >> ${P}.log
RSSM_ROLLOUT_LIB=$BASE timeout 300 python profiles/src/r2_quick.py > ${P}_base.txt 2>&1; echo "base exit $?" >> ${P}.log
timeout 300 python profiles/src/r2_quick.py > ${P}_new2.txt 2>&1
cat ${P}.log; echo NEW; cat ${P}_new.txt; echo BASE; cat ${P}_base.txt; echo NEW2; cat ${P}_new2.txt
