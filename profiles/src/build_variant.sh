#!/bin/bash
# usage: profiles/src/build_variant.sh NAME [-DFLAG ...]  -> profiles/src/lib_NAME.so
# Experiment build of the SAME ABI: the two MMTRSSM translation units are recompiled with the extra flags (default (4,2) class
# sizes only), every other object comes from the main build (multimodal_mtrssm_b200/build).  Select it with RSSM_ROLLOUT_LIB.
set -e
NAME=$1; shift
cd "$(dirname "$0")/../.."
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DRSSM_EXP_ONLY_DEFAULT $@"
nvcc $F -c multimodal_mtrssm_b200/csrc/mtrssm_kernels.cu -o profiles/src/mt_$NAME.o &
nvcc $F -c multimodal_mtrssm_b200/csrc/mtrssm_fused_bwd.cu -o profiles/src/fz_$NAME.o &
wait
B=multimodal_mtrssm_b200/build
OTHERS=$(ls $B/*.cu.o | grep -v "mtrssm_kernels\|mtrssm_fused_bwd")
nvcc -shared -o profiles/src/lib_$NAME.so profiles/src/mt_$NAME.o profiles/src/fz_$NAME.o $OTHERS
rm -f profiles/src/mt_$NAME.o profiles/src/fz_$NAME.o
echo built profiles/src/lib_$NAME.so
