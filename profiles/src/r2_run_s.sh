#!/bin/bash
# round-2 evidence run on the frozen build: all GPU tests, smoke, bench (ours + reference arm), ncu launch list and one full capture
# usage: r2_run_s.sh TAG     (files gpurun_out/TAG_*)
cd "$GRAFT_REPO_ROOT"
P=gpurun_out/${1:-s}
timeout 1200 python -m pytest tests -m gpu -q > ${P}_pytest.txt 2>&1; echo "tests exit $?" > ${P}.log
timeout 300 python __graft_entry__.py smoke > ${P}_smoke.txt 2>&1; echo "smoke exit $?" >> ${P}.log
timeout 1200 python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err; echo "bench (driver defaults) exit $?" >> ${P}.log
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 --ref-budget-s 60 > ${P}_ref.json 2> ${P}_ref.err; echo "ref exit $?" >> ${P}.log
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > ${P}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv $CMD > ${P}_ncu1.log 2>&1
echo "launch list exit $?" >> ${P}.log
$CMD > ${P}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"mtrssm_(fwd2|bwd_fused2)" -s 6 -c 2 -o ${P}_prof $CMD > ${P}_ncu2.log 2>&1
echo "full capture exit $?" >> ${P}.log
tail -3 ${P}_pytest.txt; cat ${P}.log; tail -2 ${P}_smoke.txt | cut -c1-300; tail -c 1500 ${P}_bench_default.json
