#!/bin/bash
mkdir -p gpurun_out/c53
O=gpurun_out/c53
timeout 120 python tools/unet_error.py > $O/unet_error.txt 2>&1; echo "unet_error rc=$?"; tail -1 $O/unet_error.txt
timeout 300 python -m pytest tests/test_net_gpu.py tests/test_golden_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -1 $O/pytest.log
timeout 200 python tools/layer_table.py --batch 32 > $O/layers_unet_b32.txt 2>&1
grep -E "linattn_qout  |linattn_kvctx  |forward \(|sum of ops" $O/layers_unet_b32.txt | cut -c1-70
grep -E ":linattn_qout" $O/layers_unet_b32.txt | cut -c1-40,100-130
