#!/bin/bash
# Builds libvaeseg_b200.so in-tree for sm_100a.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${VS_EXTRA_FLAGS:-} -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=default"
SRCS="cabi.cu conv3_direct.cu k2s2.cu norm_act.cu affine_act.cu losses.cu fc.cu optim.cu resize.cu"
if [ -f conv3_tc.cu ]; then SRCS="$SRCS conv3_tc.cu conv3_wgrad_tc.cu conv3_tc_kdn.cu k2s2_tc.cu k2s2_wgrad_tc.cu"; FLAGS="$FLAGS -DVS_WITH_TCGEN05"; fi
mkdir -p build
pids=()
for s in $SRCS; do
  o=build/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ vs_common.cuh -nt "$o" ] || [ tc_ptx.cuh -nt "$o" ] || [ tc_pack.cuh -nt "$o" ] || [ ../../include/vaeseg_b200.h -nt "$o" ] || [ build.sh -nt "$o" ]; then
    $NVCC $FLAGS ${VS_PTXAS_V:+-Xptxas -v} -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
OBJS=""
for s in $SRCS; do OBJS="$OBJS build/${s%.cu}.o"; done
$NVCC -shared -o libvaeseg_b200.so $OBJS -gencode arch=compute_100a,code=sm_100a -lcudart
echo "built $(pwd)/libvaeseg_b200.so"
