#!/bin/bash
# tools/build_variant.sh TAG THREADS CTAS_PER_SM [FENCE] -> pointreggpt_b200/libprg_TAG.so (load with PRG_LIB_PATH)
# Tuning builds of the fused reprojection kernel: geometry.cu recompiled with other launch constants,
# linked with the objects of the regular build.
set -e
cd "$(dirname "$0")/.."
python -m pointreggpt_b200.build > /dev/null
tag=$1; obj=/tmp/geometry_$tag.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr -DPRG_BUILDING -DPRG_RP_THREADS=$2 -DPRG_RP_CTAS_PER_SM=$3 -DPRG_RP_FENCE=${4:-1} \
  -c pointreggpt_b200/csrc/geometry.cu -o $obj
objs=$(ls pointreggpt_b200/csrc/_obj/*.o | grep -v geometry.o)
nvcc -shared -o pointreggpt_b200/libprg_$tag.so $objs $obj -lcudart
echo pointreggpt_b200/libprg_$tag.so
