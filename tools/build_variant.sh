#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG=..]...  ->  variants/libqdsim_NAME.so  (A/B builds: run with QDSIM_LIB=variants/...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o variants/libqdsim_$name.so rl-agent-for-qubit-array-tuning_b200/csrc/qd_api.cu
echo built variants/libqdsim_$name.so
