#!/bin/bash
mkdir -p gpurun_out/c23
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c23/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c23/pytest.log
tail -3 gpurun_out/c23/pytest.log
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c23/layers_unet_b32.txt 2>&1
grep -E "Utail|forward \(|sum of ops|U2:stem" gpurun_out/c23/layers_unet_b32.txt
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c23/bench_pairs.json 2> gpurun_out/c23/bench_pairs.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c23/bench_pairs.json").read().strip().splitlines()[-1])
print(d["value"], d["unet_step_ms_wall"], d["roofline"]["families_ms_per_unet_eval"])
PY
