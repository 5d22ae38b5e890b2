#!/bin/bash
# A/B build of the library with extra compile-time flags for some source files (everything else is the in-tree objects):
#   profiles/tools/build_variant.sh nolate conv_tc.cu,conv_tc2.cu -DTDVC_PDL_LATE_WAIT=0  ->  td-vc-gan_b200/tdvc/libtdvc_b200_nolate.so
# used as TDVC_LIB=td-vc-gan_b200/tdvc/libtdvc_b200_nolate.so python bench.py ...
set -e
NAME=$1; SRCS=${2//,/ }; shift 2
cd "$(dirname "$0")/../../td-vc-gan_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
objs=""; skip=""
for SRC in $SRCS; do
  $NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets "$@" -c $SRC -o /tmp/variant_${NAME}_${SRC%.cu}.o &
  objs="$objs /tmp/variant_${NAME}_${SRC%.cu}.o"; skip="$skip ${SRC%.cu}.o"
done
wait
for o in *.o; do case " $skip " in *" $o "*) ;; *) objs="$objs $o";; esac; done
$NVCC -shared -Wno-deprecated-gpu-targets -gencode arch=compute_100a,code=sm_100a -o ../tdvc/libtdvc_b200_$NAME.so $objs
echo "built td-vc-gan_b200/tdvc/libtdvc_b200_$NAME.so"
