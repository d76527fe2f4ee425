#!/bin/bash
# One GPU-box call that validates the tree and refreshes the evidence under profiles/ (about 7 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_checklist.sh'
# Outputs land in gpurun_out/checklist/; copy what should be judged into profiles/ (see profiles/README.md).
set -u
out=gpurun_out/checklist
mkdir -p "$out"
timeout 900 python -m pytest tests -q -m gpu --durations=5 > "$out/pytest_gpu.txt" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.txt"
tail -3 "$out/pytest_gpu.txt"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$out/smoke.txt" 2>&1
tail -1 "$out/smoke.txt"
timeout 600 python bench.py --steps 2 --warmup 3 > "$out/bench_pairs.json" 2> "$out/bench_pairs.err"
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > "$out/bench_geometry.json" 2> "$out/bench_geometry.err"
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 64 > "$out/bench_dataset_b4.json" 2> "$out/bench_dataset_b4.err"
for b in 32 4; do timeout 300 python tools/layer_table.py --batch $b > "$out/layer_table_unet_b$b.txt" 2>&1; done
timeout 300 python tools/layer_table.py --batch 32 --net mask > "$out/layer_table_mask_b32.txt" 2>&1
grep -E "forward \(|sum of ops" "$out"/layer_table_*.txt
# launch list of one evaluation (duration, DRAM bytes, tensor-pipe activity) -> tools/conv_traffic.py -> profiles/conv_traffic.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file "$out/launches_unet_b32.csv" python tools/profile_forward.py --batch 32 > "$out/ncu_launches.log" 2>&1
# memory checker over every kernel of an evaluation and over the geometry kernels
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/profile_forward.py --batch 2 --size 256 > "$out/memcheck_unet.log" 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_geometry.py > "$out/memcheck_geometry.log" 2>&1
grep -h "ERROR SUMMARY" "$out"/memcheck_*.log
python - <<'PY'
import json
for f in ("bench_pairs", "bench_geometry", "bench_dataset_b4"):
    try:
        d = json.loads(open("gpurun_out/checklist/%s.json" % f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, d["value"], d["unit"], "roofline.frac", r.get("frac"), "whole", r.get("whole_step_frac"), "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
