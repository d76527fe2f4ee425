#!/bin/bash
mkdir -p gpurun_out/c12
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c12/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c12/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c12/smoke.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/c12/bench_pairs.json 2> gpurun_out/c12/bench_pairs.err
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c12/bench_geometry.json 2> gpurun_out/c12/bench_geometry.err
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 64 > gpurun_out/c12/bench_dataset_b4.json 2> gpurun_out/c12/bench_dataset_b4.err
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 128 --batch 32 --no-cpu-baseline > gpurun_out/c12/bench_dataset_b32.json 2> gpurun_out/c12/bench_dataset_b32.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/c12/bench_reference.json 2> gpurun_out/c12/bench_reference.err
timeout 300 python tools/layer_table.py --batch 32 > gpurun_out/c12/layers_unet_b32.txt 2>&1
timeout 300 python tools/layer_table.py --batch 32 --net mask > gpurun_out/c12/layers_mask_b32.txt 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused|k_depth2pc_vec" -c 2 -f -o gpurun_out/c12/geom python tools/profile_geometry.py > gpurun_out/c12/ncu_geom.log 2>&1
tail -2 gpurun_out/c12/pytest.log; tail -1 gpurun_out/c12/smoke.log
python - <<'PY'
import json
for f in ["bench_pairs","bench_geometry","bench_dataset_b4","bench_dataset_b32","bench_reference"]:
    try:
        d=json.loads(open("gpurun_out/c12/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d.get("unet_step_ms_wall"), (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(f,"ERR",e)
PY
