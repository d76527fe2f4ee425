#!/bin/bash
mkdir -p gpurun_out/c28
PRG_CONV_FLAGS=512 timeout 300 python tools/unet_error.py > gpurun_out/c28/unet_error.txt 2>&1; echo "rc=$?"; tail -2 gpurun_out/c28/unet_error.txt
PRG_CONV_FLAGS=512 timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c28/layers_cg128.txt 2>&1
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c28/layers_base.txt 2>&1
grep -E "forward \(|sum of ops|conv_tc  " gpurun_out/c28/layers_cg128.txt gpurun_out/c28/layers_base.txt
grep -E "bn=128" gpurun_out/c28/layers_cg128.txt | cut -c1-60,100-130
echo ---
grep -E "bn=128" gpurun_out/c28/layers_base.txt | cut -c1-60,100-130
