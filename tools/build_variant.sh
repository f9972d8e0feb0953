#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. -DFLAG2=.."  -> variants/NAME.so (kernel tuning experiments;
# variants/ is git-ignored through *.so but travels to the GPU box, build/ does not)
set -e
cd "$(dirname "$0")/../melonix_b200/csrc"
out=../../build/variants; mkdir -p $out/$1
for f in capi multi pv_kernels pv_analyze2 spec_kernels grain_kernels picks_kernels; do
  nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $2 -c $f.cu -o $out/$1/$f.o &
done
wait
mkdir -p ../../variants
nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../../variants/$1.so $out/$1/*.o -ldl
echo built variants/$1.so
