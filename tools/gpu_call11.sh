#!/bin/bash
mkdir -p gpurun_out/c11
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c11/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c11/pytest.log
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c11/bench_geometry.json 2> gpurun_out/c11/bench_geometry.err
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c11/layers_unet_b32.txt 2>&1
PRG_NO_PDL=1 timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c11/layers_unet_b32_nopdl.txt 2>&1
timeout 300 python tools/layer_table.py --batch 4 > gpurun_out/c11/layers_unet_b4.txt 2>&1
PRG_NO_PDL=1 timeout 300 python tools/layer_table.py --batch 4 > gpurun_out/c11/layers_unet_b4_nopdl.txt 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/c11/bench_pairs.json 2> gpurun_out/c11/bench_pairs.err
PRG_NO_PDL=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c11/bench_pairs_nopdl.json 2> gpurun_out/c11/bench_pairs_nopdl.err
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 32 --no-cpu-baseline > gpurun_out/c11/bench_dataset_b4.json 2> gpurun_out/c11/bench_dataset_b4.err
tail -3 gpurun_out/c11/pytest.log; for f in gpurun_out/c11/layers*.txt; do echo $f; head -1 $f; done
python - <<'PY'
import json
for f in ["bench_geometry","bench_pairs","bench_pairs_nopdl","bench_dataset_b4"]:
    try:
        d=json.loads(open("gpurun_out/c11/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d.get("unet_step_ms_wall"), (d.get("roofline") or {}).get("reproject"))
    except Exception as e: print(f,"ERR",e)
PY
