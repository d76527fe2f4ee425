#!/bin/bash
mkdir -p gpurun_out/c35
O=gpurun_out/c35
timeout 180 python tools/unet_error.py > $O/unet_error.txt 2>&1; echo "unet_error rc=$?"; tail -1 $O/unet_error.txt
timeout 300 python tools/layer_table.py --batch 32 > $O/layers_unet_b32.txt 2>&1
grep -E "^U3:|^U5:|^U6:|^U13:|^U88:|^U100:|^U101:|forward \(|sum of ops|conv_tc  " $O/layers_unet_b32.txt | cut -c1-50,100-130
