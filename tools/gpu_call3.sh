#!/bin/bash
mkdir -p gpurun_out/c3
timeout 1200 python -m pytest tests -m gpu -q -rxXs --durations=10 > gpurun_out/c3/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3/pytest.log
timeout 120 python tools/unet_error.py > gpurun_out/c3/unet_error.txt 2>&1
PRG_CONV_FLAGS=128 timeout 120 python tools/unet_error.py >> gpurun_out/c3/unet_error.txt 2>&1
PRG_NO_XF=1 timeout 120 python tools/unet_error.py >> gpurun_out/c3/unet_error.txt 2>&1
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c3/bench_geometry.json 2> gpurun_out/c3/bench_geometry.err
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c3/layers_unet_b32.txt 2>&1
PRG_CONV_FLAGS=128 timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c3/layers_unet_b32_exact.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/c3/bench_pairs.json 2> gpurun_out/c3/bench_pairs.err
PRG_NO_XF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c3/bench_pairs_noxf.json 2> gpurun_out/c3/bench_pairs_noxf.err
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 32 --no-cpu-baseline > gpurun_out/c3/bench_dataset_b4.json 2> gpurun_out/c3/bench_dataset_b4.err
timeout 600 python bench.py --impl reference-gpu --steps 5 --warmup 3 > gpurun_out/c3/bench_refgpu.json 2> gpurun_out/c3/bench_refgpu.err
grep -n "passed\|failed" gpurun_out/c3/pytest.log | tail -3; cat gpurun_out/c3/unet_error.txt
