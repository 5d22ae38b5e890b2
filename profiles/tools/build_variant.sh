#!/bin/bash
# A/B build of the library with extra compile-time flags for ONE source file (everything else is the in-tree objects):
#   profiles/tools/build_variant.sh nohoist conv_tc.cu -DTDVC_EPI_HOIST=0   ->  td-vc-gan_b200/tdvc/libtdvc_b200_nohoist.so
# used as TDVC_LIB=td-vc-gan_b200/tdvc/libtdvc_b200_nohoist.so python bench.py ...
set -e
NAME=$1; SRC=$2; shift 2
cd "$(dirname "$0")/../../td-vc-gan_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets "$@" -c $SRC -o /tmp/variant_$NAME.o
$NVCC -shared -Wno-deprecated-gpu-targets -gencode arch=compute_100a,code=sm_100a -o ../tdvc/libtdvc_b200_$NAME.so /tmp/variant_$NAME.o $(ls *.o | grep -v "^${SRC%.cu}.o$")
echo "built td-vc-gan_b200/tdvc/libtdvc_b200_$NAME.so"
