#!/bin/bash
# A/B builds of libqmpc.so for kernel tuning (selected at run time with QMPC_LIB=...): name=flags pairs
# usage: bash scripts/build_variants.sh r2="-DQMPC_RING=2" r2p="-DQMPC_RING=2 -DQMPC_WR=18"
cd "$(dirname "$0")/../mpc_quad_ros_b200/csrc"
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $flags capi.cu -o libqmpc_$name.so 2>&1 | grep -v "warning\|^ *$\|detected during\|instantiation of\|\^" ; echo "built libqmpc_$name.so [$flags]" ) &
done
wait
