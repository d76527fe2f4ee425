#!/bin/bash
mkdir -p gpurun_out/c43
O=gpurun_out/c43
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --steps 2 --warmup 3 > $O/bench_pairs.json 2> $O/bench_pairs.err
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > $O/bench_geometry.json 2> $O/bench_geometry.err
timeout 300 python tools/layer_table.py --batch 32 > $O/layers_unet_b32.txt 2>&1
timeout 300 python tools/layer_table.py --batch 4 > $O/layers_unet_b4.txt 2>&1
timeout 300 python tools/layer_table.py --batch 32 --net mask > $O/layers_mask_b32.txt 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $O/launches_unet_b32.csv python tools/profile_forward.py --batch 32 > $O/ncu_launches.log 2>&1
tail -2 $O/pytest.log; tail -1 $O/smoke.log
grep -E "forward \(|sum of ops" $O/layers_unet_b32.txt $O/layers_unet_b4.txt $O/layers_mask_b32.txt
python - <<'PY'
import json
for f in ["bench_pairs","bench_geometry"]:
    try:
        d=json.loads(open("gpurun_out/c43/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d.get("unet_step_ms_wall"), (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("whole_step_frac"), d.get("e2e",{}).get("value"))
    except Exception as e: print(f,"ERR",e)
PY
