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
# the path library beside it (rpath $ORIGIN), so tools/path_probe.py can run the variant: HASLR_PATH_LIB=build/var/NAME/libhaslr_path.so
cp build/var/$name.so build/var/$name/libhaslr_b200.so
g++ -std=c++17 -O2 -fPIC -shared -o build/var/$name/libhaslr_path.so $(ls haslr_b200/host/*.cpp | grep -v main.cpp) -Lbuild/var/$name -lhaslr_b200 -Wl,-rpath,'$ORIGIN' -lz -lpthread
echo built build/var/$name.so
