#!/bin/bash
# ncu --set full of the fp32-parity policy's kernels at the bench batch (forward, BPTT, weight gradients: one launch each)
cd "$GRAFT_REPO_ROOT"
P=gpurun_out/${1:-nf}
CMD="python profiles/src/r2_fp32_quick.py 37888 2"
$CMD > ${P}_plain.txt 2>&1 || { echo plain run failed; tail ${P}_plain.txt; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"mtrssm_(fwd|bwd)_kernel" -s 4 -c 2 -o ${P}_prof $CMD > ${P}_ncu.log 2>&1
echo "ncu exit $?"; ls -la ${P}_prof.ncu-rep; cat ${P}_plain.txt
