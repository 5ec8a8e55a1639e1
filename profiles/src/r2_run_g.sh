#!/bin/bash
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"mtrssm_(fwd2|bwd_fused2)" -s 6 -c 2 -o gpurun_out/g_prof $CMD > gpurun_out/g_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/g_ncu.log; ls -la gpurun_out/g_prof.ncu-rep
