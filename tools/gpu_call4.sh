#!/bin/bash
mkdir -p gpurun_out/c4
timeout 1200 python -m pytest tests -m gpu -q -rxXs --durations=10 > gpurun_out/c4/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4/pytest.log
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused|k_depth2pc_vec" -c 2 -f -o gpurun_out/c4/geom python tools/profile_geometry.py > gpurun_out/c4/ncu_geom.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_conv2 --launch-skip 1 --launch-count 1 -f -o gpurun_out/c4/convxf python tools/profile_forward.py --batch 16 > gpurun_out/c4/ncu_convxf.log 2>&1
timeout 600 python tools/dual_stream.py --batch 32 --steps 30 > gpurun_out/c4/dual_stream.txt 2>&1
grep -n "passed\|failed" gpurun_out/c4/pytest.log | tail -3; cat gpurun_out/c4/dual_stream.txt; tail -3 gpurun_out/c4/ncu_geom.log; tail -3 gpurun_out/c4/ncu_convxf.log
