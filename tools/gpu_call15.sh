#!/bin/bash
mkdir -p gpurun_out/c15
timeout 600 python -m pytest tests/test_geometry_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/c15/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c15/pytest.log
tail -3 gpurun_out/c15/pytest.log
run() {
  name=$1; shift
  echo "== $name" | tee -a gpurun_out/c15/variants.txt
  env "$@" timeout 100 python tools/bench_geometry.py --only reproject --maps 500 2>&1 | tail -1 | tee -a gpurun_out/c15/variants.txt
}
run default A=1
run nodeps PRG_RP_FLAGS=2
run nofence PRG_RP_FLAGS=4
run nodeps_nofence PRG_RP_FLAGS=6
run item8k_ring48 PRG_RP_ITEM_PX=8192 PRG_RP_RING_MB=48
run item16k_ring48 PRG_RP_ITEM_PX=16384 PRG_RP_RING_MB=48
run item8k_ring24 PRG_RP_ITEM_PX=8192
run item16k_nodeps_nofence PRG_RP_ITEM_PX=16384 PRG_RP_FLAGS=6
run static PRG_RP_FLAGS=1
run maps256 A=1 
