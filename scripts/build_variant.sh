#!/bin/bash
# build_variant.sh NAME "-DQPW_UVEC=1 -DQPW_PADS=0 ..." : links ad_mpc_b200/variants/NAME.so = the in-tree objects with qp_warp.cu
# (and qp_warp_f.cu) recompiled under the given macros.  Select it at run time with ADMPC_LIB=<path>.
set -e
cd "$(dirname "$0")/../ad_mpc_b200"
mkdir -p variants
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -I csrc -I ../include"
nvcc $F $2 -c csrc/qp_warp.cu -o variants/$1.qp_warp.o
nvcc $F $2 -c csrc/qp_warp_f.cu -o variants/$1.qp_warp_f.o
OBJS=$(ls csrc/*.o | grep -v "csrc/qp_warp.o\|csrc/qp_warp_f.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o variants/$1.so $OBJS variants/$1.qp_warp.o variants/$1.qp_warp_f.o -ldl
echo variants/$1.so
