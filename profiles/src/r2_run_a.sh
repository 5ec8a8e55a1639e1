#!/bin/bash
# round-2 call A: all GPU tests (old + new parity tests at bench sizes) and a quick bench of the re-ordered fused backward
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -q -x --deselect tests/test_bench_configs_gpu.py > gpurun_out/a_pytest_old.txt 2>&1; echo "old tests exit $?" >> gpurun_out/a.log
timeout 1500 python -m pytest tests/test_bench_configs_gpu.py -q -s > gpurun_out/a_pytest_new.txt 2>&1; echo "new tests exit $?" >> gpurun_out/a.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench exit $?" >> gpurun_out/a.log
tail -3 gpurun_out/a_pytest_old.txt; tail -3 gpurun_out/a_pytest_new.txt; cut -c1-600 gpurun_out/a_bench.json
