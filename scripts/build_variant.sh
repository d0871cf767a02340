#!/bin/bash
# Build csrc/libnellie_b200_<name>.so with extra -D flags for hessian_fast.cu (A/B kernel experiments on the GPU box:
#   NB200_LIB=nellie_b200/csrc/libnellie_b200_<name>.so python scripts/prof_fast.py ...)
# usage: scripts/build_variant.sh <name> -DNB200_STATS_CTAS=3 -DNB200_STATS_D=6 ...
set -e
name=$1; shift
cd /root/repo && python -c "from nellie_b200 import build; build.build()"
cd /root/repo/nellie_b200/csrc
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC \
  --expt-relaxed-constexpr --expt-extended-lambda -c hessian_fast.cu -o /tmp/hessian_fast_$name.o
objs=$(ls *.o | grep -v hessian_fast.o)
nvcc -shared -o libnellie_b200_$name.so $objs /tmp/hessian_fast_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built libnellie_b200_$name.so
