#!/bin/bash
mkdir -p gpurun_out/c30
O=gpurun_out/c30
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_golden_gpu.py tests/test_traj_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -2 $O/pytest.log
timeout 300 python tools/layer_table.py --batch 4 > $O/layers_unet_b4.txt 2>&1
timeout 300 python tools/layer_table.py --batch 32 > $O/layers_unet_b32.txt 2>&1
grep -E "forward \(|sum of ops|linattn_kvctx  " $O/layers_unet_b4.txt $O/layers_unet_b32.txt
grep -E "kvctx" $O/layers_unet_b4.txt | cut -c1-50,100-125
