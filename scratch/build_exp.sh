#!/bin/bash
# usage: scratch/build_exp.sh NAME -DFLAG...   -> scratch/lib_NAME.so (experiment builds; MTRSSM default instantiation only)
set -e
NAME=$1; shift
cd /root/repo
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DRSSM_EXP_ONLY_DEFAULT $@"
nvcc $F -c multimodal_mtrssm_b200/csrc/mtrssm_kernels.cu -o scratch/mt_$NAME.o
nvcc -shared -o scratch/lib_$NAME.so scratch/mt_$NAME.o multimodal_mtrssm_b200/build/mrssm_kernels.cu.o multimodal_mtrssm_b200/build/wgrad_kernel.cu.o multimodal_mtrssm_b200/build/rollout_abi.cu.o
echo built scratch/lib_$NAME.so
