#!/bin/bash
mkdir -p gpurun_out/c5
timeout 1200 python -m pytest tests -m gpu -q -rxXs --durations=5 > gpurun_out/c5/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5/pytest.log
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c5/bench_geometry.json 2> gpurun_out/c5/bench_geometry.err
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c5/layers_unet_b32.txt 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused" -c 1 -f -o gpurun_out/c5/geom python tools/profile_geometry.py > gpurun_out/c5/ncu_geom.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/c5/bench_pairs.json 2> gpurun_out/c5/bench_pairs.err
grep -n "passed\|failed" gpurun_out/c5/pytest.log | tail -3; cat gpurun_out/c5/bench_geometry.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['reproject'], d['roofline']['depth2pc'])"; grep "gn_in" gpurun_out/c5/layers_unet_b32.txt | cut -c1-30,100-130; head -1 gpurun_out/c5/layers_unet_b32.txt
