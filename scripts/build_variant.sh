#!/bin/bash
# build_variant.sh NAME "-DMACRO=... ..." [file.cu ...] : links ad_mpc_b200/variants/NAME.so = the in-tree objects with the given
# sources (default: qp_mma.cu) recompiled under the given macros.  Select it at run time with ADMPC_LIB=<path>.
set -e
cd "$(dirname "$0")/../ad_mpc_b200"
mkdir -p variants
NAME=$1; MACROS=$2; shift 2 || true
FILES=${@:-qp_mma.cu}
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -I csrc -I ../include -Xptxas -v"
EXCL=""
OBJS=""
for f in $FILES; do
  b=$(basename $f .cu)
  nvcc $F $MACROS -c csrc/$b.cu -o variants/$NAME.$b.o 2> variants/$NAME.$b.log
  OBJS="$OBJS variants/$NAME.$b.o"
  EXCL="$EXCL|csrc/$b.o"
done
REST=$(ls csrc/*.o | grep -v -E "^(${EXCL:1})$" || true)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o variants/$NAME.so $REST $OBJS -ldl
grep -h "registers\|spill" variants/$NAME.*.log | paste - - | sed 's/ptxas info    ://g' | head -4
echo variants/$NAME.so
