#!/bin/bash
# usage: profiles/src/build_fz.sh NAME -DFLAG...  -> profiles/src/lib_NAME.so (fused backward experiment builds, default instantiation only)
set -e
NAME=$1; shift
cd /root/repo
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DRSSM_EXP_ONLY_DEFAULT $@"
nvcc $F -c multimodal_mtrssm_b200/csrc/mtrssm_fused_bwd.cu -o profiles/src/fz_$NAME.o
B=multimodal_mtrssm_b200/build
nvcc -shared -o profiles/src/lib_$NAME.so profiles/src/fz_$NAME.o $B/mrssm_kernels.cu.o $B/mtrssm_kernels.cu.o $B/wgrad_kernel.cu.o $B/rollout_abi.cu.o
echo built profiles/src/lib_$NAME.so
