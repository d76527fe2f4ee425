#!/bin/bash
mkdir -p gpurun_out/c32
O=gpurun_out/c32
timeout 180 python tools/unet_error.py > $O/unet_error.txt 2>&1; echo "unet_error rc=$?"; tail -2 $O/unet_error.txt
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_net_gpu.py tests/test_golden_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python tools/layer_table.py --batch 32 > $O/layers_unet_b32.txt 2>&1
PRG_CONV_FLAGS=256 timeout 300 python tools/layer_table.py --batch 32 > $O/layers_unet_b32_nodxs.txt 2>&1
grep -E "forward \(|sum of ops|conv_tc  " $O/layers_unet_b32.txt $O/layers_unet_b32_nodxs.txt
grep -E "dxs=1" $O/layers_unet_b32.txt | cut -c1-60,100-130
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_traj_gpu.py -m gpu -q -x -s > $O/pytest_s.log 2>&1; echo "pytest rc=$?" >> $O/pytest_s.log
grep -E "fused vs separate|passed|failed|rc=" $O/pytest_s.log | tail -5
