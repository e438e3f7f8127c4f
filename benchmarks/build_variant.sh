#!/bin/bash
# build_variant.sh NAME SRC.cu FLAGS... : compile one TU of csrc/ with extra -D flags and link it with the regular objects
# into ndarray_b200/variants/libndb200_NAME.so (select it with NDB200_LIB=...). Kernel experiments only.
set -e
root=/root/repo/ndarray_b200
name=$1; src=$2; shift 2
mkdir -p $root/variants
base=$(basename $src .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC --expt-relaxed-constexpr -DNDB_BUILDING "$@" -c $root/csrc/$src -o $root/variants/${base}_$name.o
objs=$(ls $root/build/*.o | grep -v "/$base.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $root/variants/libndb200_$name.so $objs $root/variants/${base}_$name.o \
  -Xcompiler -fPIC -cudart static
echo built $root/variants/libndb200_$name.so
