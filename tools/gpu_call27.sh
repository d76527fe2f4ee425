#!/bin/bash
mkdir -p gpurun_out/c27
timeout 300 python tools/unet_error.py > gpurun_out/c27/unet_error.txt 2>&1; echo "rc=$?"; tail -2 gpurun_out/c27/unet_error.txt
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_golden_gpu.py tests/test_conv_gpu.py -m gpu -q -x > gpurun_out/c27/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c27/pytest.log
tail -3 gpurun_out/c27/pytest.log
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c27/layers_unet_b32.txt 2>&1
grep -E "rows3=1|forward \(|sum of ops|conv_tc  " gpurun_out/c27/layers_unet_b32.txt
