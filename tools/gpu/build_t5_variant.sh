#!/bin/bash
# build_t5_variant.sh NAME [-D...]  -> variants/NAME.so : libackb with ackb_ppo_tcgen05.cu rebuilt under extra flags (tuning only; the other
# objects come from build/, i.e. run __graft_entry__.build() first)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
name=$1; shift
mkdir -p $ROOT/variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ftz=true -diag-suppress 177"
/usr/local/cuda/bin/nvcc $F -I$ROOT/include -I$ROOT/mujoco_playground_b200/csrc -Xcompiler -fPIC "$@" -c -o $ROOT/variants/$name.o $ROOT/mujoco_playground_b200/csrc/ackb_ppo_tcgen05.cu
/usr/local/cuda/bin/nvcc $F -shared -o $ROOT/variants/$name.so $ROOT/build/ackb_kernels.o $ROOT/build/ackb_ppo.o $ROOT/variants/$name.o -ldl
rm -f $ROOT/variants/$name.o
echo built variants/$name.so
