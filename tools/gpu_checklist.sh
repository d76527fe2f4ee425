#!/bin/bash
# One GPU-box call that answers everything open at the start of a round (about 8 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_checklist.sh'
# Outputs land in gpurun_out/checklist/.  XPASS lines in pytest_gpu.txt mean a staged kernel
# (tests/test_zz_staged_gpu.py) is validated and its xfail marker can go.
set -u
out=gpurun_out/checklist
mkdir -p "$out"
timeout 420 python -m pytest tests -q -m gpu -rxX > "$out/pytest_gpu.txt" 2>&1
tail -15 "$out/pytest_gpu.txt"
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$out/smoke.txt" 2>&1
tail -2 "$out/smoke.txt"
timeout 120 python tools/layer_table.py --batch 32 > "$out/layer_table_unet_b32.txt" 2>&1
tail -12 "$out/layer_table_unet_b32.txt"
timeout 60 python tools/bench_geometry.py > "$out/bench_geometry.txt" 2>&1
tail -2 "$out/bench_geometry.txt"
timeout 420 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"
tail -1 "$out/bench_n1.json"
