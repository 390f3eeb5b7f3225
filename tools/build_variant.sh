#!/bin/sh
# Build a LOCO_DEBUG library variant with extra nvcc flags into tools/_libs/<name>.so:   tools/build_variant.sh <name> [-DFLAG ...]
set -e
cd "$(dirname "$0")/../loco_asr_b200/csrc"
name=$1; shift
out=../../tools/_libs; tmp=/tmp/loco_variant_$name
mkdir -p $out $tmp
for f in api tensormap gemm_tcgen05_2cta frontend conv0_tc conv0_mma rowops posconv_pp posconv_tc attention_tc attention_p2 gemm_tcgen05 gemm_simt posconv attention; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DLOCO_DEBUG "$@" -c $f.cu -o $tmp/$f.o 2>/dev/null &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $tmp/*.o -cudart static
echo built $out/$name.so
