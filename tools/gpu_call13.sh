#!/bin/bash
mkdir -p gpurun_out/c13
for mb in 8 16 24 48 96; do
  PRG_RP_RING_MB=$mb timeout 200 python bench.py --workload geometry --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/c13/geo_ring$mb.json 2> gpurun_out/c13/geo_ring$mb.err
  python - $mb <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/c13/geo_ring%s.json"%sys.argv[1]).read().strip().splitlines()[-1]); print("ring MB",sys.argv[1], d["roofline"]["reproject"])
except Exception as e: print("ERR",e)
PY
done
