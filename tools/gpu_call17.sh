#!/bin/bash
mkdir -p gpurun_out/c17
L=$PWD/pointreggpt_b200
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused" -c 1 -f -o gpurun_out/c17/rp_a python tools/profile_geometry.py > gpurun_out/c17/ncu_a.log 2>&1
PRG_LIB_PATH=$L/libprg_b.so PRG_RP_ITEM_PX=8192 PRG_RP_RING_MB=48 timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused" -c 1 -f -o gpurun_out/c17/rp_b python tools/profile_geometry.py > gpurun_out/c17/ncu_b.log 2>&1
tail -3 gpurun_out/c17/ncu_a.log gpurun_out/c17/ncu_b.log
