#!/bin/bash
# Builds libtdvc_b200.so (sm_100a only) in-tree: td-vc-gan_b200/tdvc/libtdvc_b200.so
set -e
cd "$(dirname "$0")/td-vc-gan_b200/csrc"
mkdir -p ../tdvc
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets"
objs=""
pids=""
for f in *.cu; do
  o="${f%.cu}.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ tc_common.cuh -nt "$o" ] || [ ../../include/tdvc_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    # compile to a temporary name: an interrupted or failed compile must not leave a fresh-looking object behind
    ( $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$f" -o "$o.tmp" && mv "$o.tmp" "$o" ) &
    pids="$pids $!"
  fi
  objs="$objs $o"
done
for p in $pids; do wait $p; done      # set -e: a failed compile stops the build here
$NVCC -shared -Wno-deprecated-gpu-targets -gencode arch=compute_100a,code=sm_100a -o ../tdvc/libtdvc_b200.so $objs
echo "built td-vc-gan_b200/tdvc/libtdvc_b200.so"
