#!/bin/bash
# Build libssn_b200.so in-tree for sm_100a (the .so is git-ignored but travels with gpurun).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -I../../include"
mkdir -p build
pids=()
for f in engine simt_conv simt_glue heads umma_conv umma_conv_v2 umma_wgrad s2d_glue glue_fp16 tc_glue glue_fp32 detect bn_train; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ umma_conv.cuh -nt build/$f.o ] || [ umma_dev.cuh -nt build/$f.o ] || [ umma_epi32.cuh -nt build/$f.o ] || [ ../../include/ssnb.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../ssn_b200/libssn_b200.so build/engine.o build/simt_conv.o build/simt_glue.o build/heads.o build/umma_conv.o build/umma_conv_v2.o build/umma_wgrad.o build/s2d_glue.o build/glue_fp16.o build/tc_glue.o build/glue_fp32.o build/detect.o build/bn_train.o -lcudart_static -ldl -lpthread -lrt
echo "built ../ssn_b200/libssn_b200.so"
