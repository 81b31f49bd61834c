#!/bin/bash
# usage: build/variant.sh NAME -DFLAG=... : builds build/libhp_NAME.so from hiphase_b200/csrc with extra flags
set -e
cd "$(dirname "$0")/../hiphase_b200/csrc"; mkdir -p ../../build
name=$1; shift
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -shared -o ../../build/libhp_$name.so astar_kernels.cu hp_api.cu wfa_kernels.cu post_kernels.cu local_kernels.cu assemble_kernels.cu hp_pack.cu hp_shard.cu hp_realign.cu hp_service.cu -lcudart -ldl
