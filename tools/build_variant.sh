#!/bin/bash
# Development aid: build libmultiexp with extra nvcc defines into build/variants/<name>.so
# usage: tools/build_variant.sh name -DFOO=1 ...
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants/$name
for f in msm msm_bn254 msm_secp abi abi_secp pint multi lat; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -diag-suppress 128 "$@" -c porla_b200/csrc/$f.cu -o build/variants/$name/$f.o 2>/dev/null &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so build/variants/$name/*.o -lpthread && echo built build/variants/$name.so
