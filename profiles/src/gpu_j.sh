#!/bin/bash
# full GPU suite + likelihood profile (profiles/r1_j_*)
python -m pytest tests -m gpu -q > gpurun_out/j_pytest.txt 2>&1; tail -3 gpurun_out/j_pytest.txt
python profiles/src/nll_bench.py 4096 30 > gpurun_out/j_nll_bench.json 2> gpurun_out/j_nll_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/j_nll_launches.csv python profiles/src/nll_bench.py 4096 30 2 > gpurun_out/j_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gaussian_nll -s 4 -c 2 -f -o gpurun_out/j_nll_prof python profiles/src/nll_bench.py 4096 30 2 > gpurun_out/j_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gaussian_nll_bwd -s 2 -c 1 -f -o gpurun_out/j_nll_prof_bwd python profiles/src/nll_bench.py 4096 30 2 > gpurun_out/j_ncu3.log 2>&1
python bench.py --steps 30 > gpurun_out/j_bench_full.json 2> gpurun_out/j_bench_full.err; tail -c 300 gpurun_out/j_bench_full.json
