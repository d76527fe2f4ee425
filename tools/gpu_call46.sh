#!/bin/bash
mkdir -p gpurun_out/c46
O=gpurun_out/c46
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_driver_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 64 --no-cpu-baseline > $O/bench_dataset_b4.json 2> $O/bench_dataset_b4.err
timeout 600 python bench.py --workload dataset --steps 1 --warmup 1 --pairs 64 --device-batch 4 --no-cpu-baseline > $O/bench_dataset_b4_literal.json 2> $O/bench_dataset_b4_literal.err
python - <<'PY'
import json
for f in ["bench_dataset_b4","bench_dataset_b4_literal"]:
    try:
        d=json.loads(open("gpurun_out/c46/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["config"]["batch_size"], d["config"]["device_batch"])
    except Exception as e: print(f,"ERR",e)
PY
