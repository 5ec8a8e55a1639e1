#!/bin/bash
# ncu --set full of the wide-family (cfg3) kernels: one launch each of forward, row-statistics, BPTT, d_embed, weight gradients
cd "$GRAFT_REPO_ROOT"
P=gpurun_out/${1:-nw}
CMD="python profiles/src/wide_bench.py 1024 64 grad"
$CMD > ${P}_plain.txt 2>&1 || { echo plain run failed; tail ${P}_plain.txt; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"mrssm_wide_(fwd|bwd)_kernel" -s 4 -c 2 -o ${P}_prof $CMD > ${P}_ncu.log 2>&1
echo "ncu exit $?"; ls -la ${P}_prof.ncu-rep; cat ${P}_plain.txt
