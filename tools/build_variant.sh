#!/bin/bash
# tools/build_variant.sh NAME FILE.cu "FLAGS": liborbgpu with ONE translation unit rebuilt with other -D flags, as gpurun_variants/lib_NAME.so
# (git-ignored, travels with gpurun snapshots; select with ORBGPU_LIB=gpurun_variants/lib_NAME.so).  Tuning helper only.
set -e
cd "$(dirname "$0")/../refactored_orb_slam2_b200/csrc"
mkdir -p ../../gpurun_variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC $ARCH -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC $3 -c $2 -o /tmp/variant_$1.o
OBJS=$(ls *.o | grep -v "^${2%.cu}.o$")
$NVCC $ARCH -shared -o ../../gpurun_variants/lib_$1.so $OBJS /tmp/variant_$1.o -lcudart_static -lpthread -ldl -lrt
echo built gpurun_variants/lib_$1.so
