#!/bin/bash
# one ncu --set full capture of the two rollout kernels at the bench batch (after a plain run of the same command exited 0)
# usage: r2_run_ncu.sh TAG     (files gpurun_out/TAG_*)
cd "$GRAFT_REPO_ROOT"
P=gpurun_out/${1:-ncu}
CMD="python profiles/src/r2_quick.py --only-bench"
$CMD > ${P}_plain.txt 2>&1 || { echo plain run failed; tail ${P}_plain.txt; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"mtrssm_(fwd2|bwd_fused2)" -s 8 -c 2 -o ${P}_prof $CMD > ${P}_ncu.log 2>&1
echo "ncu exit $?"; ls -la ${P}_prof.ncu-rep; cat ${P}_plain.txt
