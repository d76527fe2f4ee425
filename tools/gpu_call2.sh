#!/bin/bash
# round 2, GPU call 2: full strict suite (fused reprojection, sampler graph, trajectory parity), benches
mkdir -p gpurun_out/c2
timeout 1200 python -m pytest tests -m gpu -q -rxXs --durations=10 > gpurun_out/c2/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2/pytest.log
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c2/bench_geometry.json 2> gpurun_out/c2/bench_geometry.err
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c2/layers_unet_b32.txt 2>&1
PRG_NO_XF=1 timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c2/layers_unet_b32_noxf.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/c2/bench_pairs.json 2> gpurun_out/c2/bench_pairs.err
PRG_NO_GRAPH=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c2/bench_pairs_nograph.json 2> gpurun_out/c2/bench_pairs_nograph.err
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 32 --no-cpu-baseline > gpurun_out/c2/bench_dataset_b4.json 2> gpurun_out/c2/bench_dataset_b4.err
PRG_NO_GRAPH=1 timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 32 --no-cpu-baseline > gpurun_out/c2/bench_dataset_b4_nograph.json 2> gpurun_out/c2/bench_dataset_b4_nograph.err
timeout 600 python bench.py --impl reference-gpu --steps 5 --warmup 3 > gpurun_out/c2/bench_refgpu.json 2> gpurun_out/c2/bench_refgpu.err
tail -25 gpurun_out/c2/pytest.log
