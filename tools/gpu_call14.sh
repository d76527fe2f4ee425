#!/bin/bash
mkdir -p gpurun_out/c14
timeout 600 python -m pytest tests/test_geometry_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/c14/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c14/pytest.log
tail -3 gpurun_out/c14/pytest.log
run() {
  name=$1; shift
  env "$@" timeout 200 python bench.py --workload geometry --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/c14/geo_$name.json 2> gpurun_out/c14/geo_$name.err
  python - $name <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/c14/geo_%s.json"%sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["roofline"]["reproject"], d["checked"])
except Exception as e: print("ERR",e)
PY
}
run static PRG_RP_STATIC=1
run dyn1 PRG_RP_AHEAD=1
run dyn2 PRG_RP_AHEAD=2
run dyn1_ring48 PRG_RP_AHEAD=1 PRG_RP_RING_MB=48
run dyn1_ring12 PRG_RP_AHEAD=1 PRG_RP_RING_MB=12
