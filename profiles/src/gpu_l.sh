#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/n_pytest.txt 2>&1; tail -3 gpurun_out/l_pytest.txt
python __graft_entry__.py smoke > gpurun_out/n_smoke.txt 2>&1; tail -1 gpurun_out/l_smoke.txt | cut -c 1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/n_bench_reference_arm.json 2> gpurun_out/n_ref.err
python bench.py > gpurun_out/n_bench_full.json 2> gpurun_out/n_bench_full.err; tail -c 200 gpurun_out/n_bench_full.err
python bench.py --no-extras > gpurun_out/n_bench_default.json 2> gpurun_out/n_bench_default.err
