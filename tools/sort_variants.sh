#!/bin/bash
# Builds libcpm_b200 variants with different onesweep tile shapes into tools/variants/ (here, no GPU needed);
# `tools/sort_variants.sh run` (on the GPU box) times each with tools/quickbench.py sort26.
PKG=correlated-photon-mapping-for-interactive-global-illumination-of-time-varying-volumetric-data_b200
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -Iinclude -I$PKG/csrc"
VARIANTS="512:16:2:4 384:16:3:4 256:16:4:4 512:16:2:8 512:16:2:2 384:20:2:4 256:24:3:4 1024:8:1:4"
if [ "$1" = "run" ]; then
  for v in $VARIANTS; do
    echo "== variant threads:items:minblocks:lookback = $v"
    CPM_B200_LIB=$PWD/tools/variants/libcpm_b200_${v//:/_}.so timeout 120 python tools/quickbench.py sort26 2>&1 | grep "sort k"
  done
  exit 0
fi
mkdir -p tools/variants
for v in $VARIANTS; do
  IFS=: read T I B W <<< "$v"
  $NV -DCPM_SORT_THREADS=$T -DCPM_SORT_ITEMS=$I -DCPM_SORT_MIN_BLOCKS=$B -DCPM_SORT_LOOKBACK=$W -Xptxas -v -c $PKG/csrc/radixsort.cu -o tools/variants/rs_${v//:/_}.o 2>&1 | grep -A1 "onesweep" | grep -E "spill" | tr '\n' ' '
  echo " <- $v"
  OBJS=$(ls build/*.o | grep -v radixsort)
  /usr/local/cuda/bin/nvcc -shared -o tools/variants/libcpm_b200_${v//:/_}.so $OBJS tools/variants/rs_${v//:/_}.o -gencode arch=compute_100a,code=sm_100a -cudart static -ldl
done
