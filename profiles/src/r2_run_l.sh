#!/bin/bash
# round-2 evidence run on the frozen build: all GPU tests, smoke, bench (ours + reference arm), ncu launch list and one full capture
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.txt 2>&1; echo "tests exit $?" > gpurun_out/l.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/l_smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/l.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench exit $?" >> gpurun_out/l.log
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 --ref-budget-s 60 > gpurun_out/l_ref.json 2> gpurun_out/l_ref.err; echo "ref exit $?" >> gpurun_out/l.log
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/l_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l_launches.csv $CMD > gpurun_out/l_ncu1.log 2>&1
echo "launch list exit $?" >> gpurun_out/l.log
$CMD > gpurun_out/l_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"mtrssm_(fwd2|bwd_fused2)" -s 6 -c 2 -o gpurun_out/l_prof $CMD > gpurun_out/l_ncu2.log 2>&1
echo "full capture exit $?" >> gpurun_out/l.log
tail -3 gpurun_out/l_pytest.txt; cat gpurun_out/l.log; tail -2 gpurun_out/l_smoke.txt | cut -c1-300; tail -c 1500 gpurun_out/l_bench.json
