#!/bin/bash
# dev tool: build an experimental variant of the library:  tools/build_variant.sh <name> <extra nvcc flags...>
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/var_$name
for f in kernels gram gram_fused capi; do
  if [ "$f" = "kernels" ] && [ -f build/kernels.o ]; then cp build/kernels.o build/var_$name/kernels.o; continue; fi
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ "$@" -c rosdyn_b200/csrc/$f.cu -o build/var_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -cudart static -ccbin /usr/bin/g++ -o build/var_$name/librosdyn_b200.so build/var_$name/*.o
echo build/var_$name/librosdyn_b200.so
