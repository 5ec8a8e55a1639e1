#!/bin/bash
# fp32-parity policy: all GPU parity tests, then kernel timings at the bench batch
cd "$GRAFT_REPO_ROOT"
P=gpurun_out/${1:-fp}
timeout 900 python -m pytest tests -m gpu -q -x > ${P}_pytest.txt 2>&1; echo "tests exit $?" > ${P}.log
tail -15 ${P}_pytest.txt
timeout 300 python profiles/src/r2_fp32_quick.py 37888 5 > ${P}_fp32.txt 2>&1; cat ${P}.log ${P}_fp32.txt
