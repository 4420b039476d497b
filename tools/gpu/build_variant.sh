#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...]  -> build/variants/NAME.so   (tuning only)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
name=$1; shift
mkdir -p $ROOT/build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ftz=true -diag-suppress 177 \
  -I$ROOT/include -I$ROOT/mujoco_playground_b200/csrc -shared -Xcompiler -fPIC "$@" \
  -o $ROOT/build/variants/$name.so $ROOT/mujoco_playground_b200/csrc/ackb_kernels.cu $ROOT/mujoco_playground_b200/csrc/ackb_ppo.cu
echo built $name
