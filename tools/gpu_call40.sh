#!/bin/bash
mkdir -p gpurun_out/c40
O=gpurun_out/c40
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_res1x1_gn" -c 5 -f -o $O/resgn python tools/profile_forward.py --batch 32 > $O/ncu_resgn.log 2>&1
# conv launches 42 .. 48 of the evaluation: dxs x 4, rows3, two-source halo, input-transform conv
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_conv2" --launch-skip 42 --launch-count 7 -f -o $O/conv_sel python tools/profile_forward.py --batch 32 > $O/ncu_conv.log 2>&1
tail -2 $O/ncu_resgn.log $O/ncu_conv.log
ls -la $O
