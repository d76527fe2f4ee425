#!/bin/bash
mkdir -p gpurun_out/c26
timeout 300 python tools/unet_error.py > gpurun_out/c26/unet_error.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c26/unet_error.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c26/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c26/pytest.log
tail -5 gpurun_out/c26/pytest.log
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c26/layers_unet_b32.txt 2>&1
grep -E "c4 \[|forward \(|sum of ops|conv_tc  " gpurun_out/c26/layers_unet_b32.txt
PRG_NO_ROWS3=1 timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c26/layers_unet_b32_norows3.txt 2>&1
grep -E "forward \(|sum of ops" gpurun_out/c26/layers_unet_b32_norows3.txt
timeout 300 python tools/layer_table.py --batch 32 --net mask > gpurun_out/c26/layers_mask_b32.txt 2>&1
grep -E "forward \(|sum of ops" gpurun_out/c26/layers_mask_b32.txt
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c26/bench_pairs.json 2> gpurun_out/c26/bench_pairs.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c26/bench_pairs.json").read().strip().splitlines()[-1])
print(d["value"], d["unet_step_ms_wall"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["roofline"]["families_ms_per_unet_eval"])
PY
