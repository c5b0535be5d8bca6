#!/bin/bash
# profiling build of the library with the tail probe compiled in (never the product library): gym_puzzles_b200/csrc/libmrp_probe.so
set -e
cd "$(dirname "$0")/../gym_puzzles_b200/csrc"
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -I ../../include -DMRP_TAILPROBE"
mkdir -p build
nvcc $F -c -o build/p_mrp_b200.o mrp_b200.cu &
nvcc $F -DMRP_MAXC=192 -c -o build/p_mrp_b200_wide.o mrp_b200.cu &
nvcc $F -c -o build/p_mrp_vecnorm.o mrp_vecnorm.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libmrp_probe.so build/p_mrp_b200.o build/p_mrp_b200_wide.o build/p_mrp_vecnorm.o
