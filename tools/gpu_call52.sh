#!/bin/bash
mkdir -p gpurun_out/c52
O=gpurun_out/c52
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 400 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_pairs.json 2> $O/bench_pairs.err
timeout 200 python tools/layer_table.py --batch 32 > $O/layers_unet_b32.txt 2>&1
tail -2 $O/pytest.log; tail -1 $O/smoke.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c52/bench_pairs.json").read().strip().splitlines()[-1])
print(d["value"], d.get("unet_step_ms_wall"), d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["e2e"]["value"])
print(d["roofline"]["families_ms_per_unet_eval"])
PY
