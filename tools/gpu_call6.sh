#!/bin/bash
mkdir -p gpurun_out/c6
timeout 1200 python -m pytest tests -m gpu -q -rxXs --durations=5 > gpurun_out/c6/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6/pytest.log
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c6/bench_geometry.json 2> gpurun_out/c6/bench_geometry.err
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c6/layers_unet_b32.txt 2>&1
PRG_GNRES=256 timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c6/layers_unet_b32_gnres256.txt 2>&1
PRG_GNRES=1 timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c6/layers_unet_b32_gnres.txt 2>&1
PRG_GNRES=1 timeout 120 python tools/unet_error.py > gpurun_out/c6/unet_error.txt 2>&1
timeout 120 python tools/unet_error.py >> gpurun_out/c6/unet_error.txt 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused" -c 1 -f -o gpurun_out/c6/geom python tools/profile_geometry.py > gpurun_out/c6/ncu_geom.log 2>&1
grep -n "passed\|failed" gpurun_out/c6/pytest.log | tail -3; cat gpurun_out/c6/bench_geometry.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['reproject'], d['roofline']['depth2pc'])"; for f in gpurun_out/c6/layers_unet_b32*.txt; do head -1 $f; done; grep "gn_in" gpurun_out/c6/layers_unet_b32.txt | cut -c1-30,100-130 | head -3; cat gpurun_out/c6/unet_error.txt
