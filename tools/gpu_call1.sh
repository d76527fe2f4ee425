#!/bin/bash
# round 2, GPU call 1: strict test suite, z-buffer micro-benchmark, baseline benches and layer tables
mkdir -p gpurun_out/c1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/c1/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rxXs -x --durations=15 > gpurun_out/c1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1/pytest.log
timeout 120 tools/microbench/bin/zbuf_atomics > gpurun_out/c1/zbuf_atomics.txt 2>&1
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c1/layers_unet_b32.txt 2>&1
timeout 300 python tools/layer_table.py --batch 4 > gpurun_out/c1/layers_unet_b4.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/c1/bench_pairs.json 2> gpurun_out/c1/bench_pairs.err
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c1/bench_geometry.json 2> gpurun_out/c1/bench_geometry.err
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 32 > gpurun_out/c1/bench_dataset_b4.json 2> gpurun_out/c1/bench_dataset_b4.err
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 64 --batch 32 --no-cpu-baseline > gpurun_out/c1/bench_dataset_b32.json 2> gpurun_out/c1/bench_dataset_b32.err
timeout 600 python bench.py --impl reference-gpu --steps 5 --warmup 3 > gpurun_out/c1/bench_refgpu.json 2> gpurun_out/c1/bench_refgpu.err
tail -5 gpurun_out/c1/pytest.log
