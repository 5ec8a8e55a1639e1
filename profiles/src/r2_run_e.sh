#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/e_pytest.txt 2>&1; echo "tests exit $?" > gpurun_out/e.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench exit $?" >> gpurun_out/e.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 40 > gpurun_out/e_ref.json 2> gpurun_out/e_ref.err; echo "ref exit $?" >> gpurun_out/e.log
tail -4 gpurun_out/e_pytest.txt; cat gpurun_out/e.log; tail -c 1800 gpurun_out/e_bench.json; echo; tail -3 gpurun_out/e_bench.err; cut -c1-400 gpurun_out/e_ref.json
