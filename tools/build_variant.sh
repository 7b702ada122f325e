#!/bin/bash
# A/B builds: tools/build_variant.sh NAME FILE.cu "-DFLAG=.. ..." -> abvariants/NAME/{libcpm_b200.so,libcpm_host.so}
# (FILE.cu recompiled with the flags, every other object taken from build/).  Select at run time with
#   CPM_B200_LIB=abvariants/NAME/libcpm_b200.so CPM_HOST_LIB=abvariants/NAME/libcpm_host.so python bench.py ...
set -e
NAME=$1; FILE=$2; FLAGS=$3
PKG=correlated-photon-mapping-for-interactive-global-illumination-of-time-varying-volumetric-data_b200
mkdir -p abvariants/$NAME build/variant_$NAME
NCCL_INC=
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
    -Xcompiler -fPIC,-fvisibility=hidden,-Wall -Iinclude -I$PKG/csrc ${NCCL_INC:+-I$NCCL_INC} $FLAGS -c $PKG/csrc/$FILE -o build/variant_$NAME/${FILE%.cu}.o
OBJS=$(ls build/*.o | grep -v "/${FILE%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -o abvariants/$NAME/libcpm_b200.so $OBJS build/variant_$NAME/${FILE%.cu}.o -gencode arch=compute_100a,code=sm_100a -cudart static -ldl
cp $PKG/libcpm_host.so abvariants/$NAME/libcpm_host.so
cuobjdump -res-usage build/variant_$NAME/${FILE%.cu}.o 2>/dev/null | grep -A1 "${VARIANT_GREP:-trace_kernelILi2ELi1ELi[12]}" | grep -o "REG:[0-9]*" | tr '\n' ' '; echo " <- $NAME"
