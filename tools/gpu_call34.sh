#!/bin/bash
mkdir -p gpurun_out/c34
O=gpurun_out/c34
PRG_CONV_FLAGS=1024 timeout 180 python tools/unet_error.py > $O/unet_error.txt 2>&1; echo "unet_error rc=$?"; tail -1 $O/unet_error.txt
PRG_CONV_FLAGS=1024 timeout 300 python tools/layer_table.py --batch 32 > $O/layers_wres.txt 2>&1
timeout 300 python tools/layer_table.py --batch 32 > $O/layers_base.txt 2>&1
grep -E "k4 m1|forward \(" $O/layers_wres.txt | cut -c1-70,100-130
grep -E "k4 m1|forward \(" $O/layers_base.txt | cut -c1-70,100-130
