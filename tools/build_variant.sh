#!/bin/bash
# tools/build_variant.sh NAME "-DBJ_WARPS=4 ..." -- a second build of the library with extra nvcc
# defines, as guido_b200/lib/variants/NAME.so (select it with GUIDO_B200_LIB=<path>); tuning only.
set -e
cd "$(dirname "$0")/../guido_b200/csrc"
name=$1; shift
out=../lib/variants; mkdir -p $out/obj_$name
for f in capi pam_scan offtarget join_blocks mmej sort; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c $f.cu -o $out/obj_$name/$f.o &
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -x cu -c image.cpp -o $out/obj_$name/image.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $out/obj_$name/*.o
rm -rf $out/obj_$name
echo built $out/$name.so
