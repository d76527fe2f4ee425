#!/bin/bash
mkdir -p gpurun_out/c10
timeout 600 python -m pytest tests/test_geometry_gpu.py -m gpu -q > gpurun_out/c10/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10/pytest.log
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c10/bench_geometry.json 2> gpurun_out/c10/bench_geometry.err
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused" -c 1 -f -o gpurun_out/c10/geom python tools/profile_geometry.py > gpurun_out/c10/ncu_geom.log 2>&1
tail -2 gpurun_out/c10/pytest.log; cat gpurun_out/c10/bench_geometry.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['reproject'], d['roofline']['depth2pc'])"
