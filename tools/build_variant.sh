#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG=..]...   ->  build/var/NAME.so  (same sources, extra defines; for A/B runs with HASLR_B200_LIB)
set -e
name=$1; shift
mkdir -p build/var/$name
for t in api poa k12 coords paf; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function "$@" -c haslr_b200/csrc/$t.cu -o build/var/$name/$t.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/var/$name.so build/var/$name/*.o -lcudart
echo built build/var/$name.so
