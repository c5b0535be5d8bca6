#!/bin/bash
# side builds of the library (never the product library):
#   bash profiles/build_probe.sh          tail probe compiled in   -> gym_puzzles_b200/csrc/libmrp_probe.so
#   bash profiles/build_probe.sh check    in-kernel bounds checks  -> gym_puzzles_b200/csrc/libmrp_check.so
set -e
cd "$(dirname "$0")/../gym_puzzles_b200/csrc"
DEF=-DMRP_TAILPROBE; OUT=libmrp_probe.so
if [ "$1" = "check" ]; then DEF=-DMRP_CHECK; OUT=libmrp_check.so; fi
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -I ../../include $DEF"
mkdir -p build
nvcc $F -c -o build/p_mrp_b200.o mrp_b200.cu &
nvcc $F -DMRP_MAXC=192 -c -o build/p_mrp_b200_wide.o mrp_b200.cu &
nvcc $F -c -o build/p_mrp_vecnorm.o mrp_vecnorm.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/p_mrp_b200.o build/p_mrp_b200_wide.o build/p_mrp_vecnorm.o
