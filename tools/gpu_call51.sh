#!/bin/bash
mkdir -p gpurun_out/c51
O=gpurun_out/c51
timeout 180 python tools/unet_error.py > $O/unet_error.txt 2>&1; echo "unet_error rc=$?"; tail -1 $O/unet_error.txt
timeout 600 python -m pytest tests/test_net_gpu.py tests/test_golden_gpu.py tests/test_conv_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -2 $O/pytest.log
timeout 300 python tools/layer_table.py --batch 32 > $O/layers_unet_b32.txt 2>&1
grep -E "^U3:|^U5:|^U47:|^U88:|^U98:|^U99:|forward \(|sum of ops|conv_tc  " $O/layers_unet_b32.txt | cut -c1-50,100-150
