#!/bin/bash
mkdir -p gpurun_out/c18
timeout 600 python -m pytest tests/test_geometry_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/c18/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c18/pytest.log
tail -3 gpurun_out/c18/pytest.log
run() {
  name=$1; shift
  echo "== $name" | tee -a gpurun_out/c18/variants.txt
  env "$@" timeout 100 python tools/bench_geometry.py --only reproject --maps 500 2>&1 | tail -1 | tee -a gpurun_out/c18/variants.txt
}
L=$PWD/pointreggpt_b200
run a_256x3 A=1
run a_256x3_r48 PRG_RP_RING_MB=48
run a_256x3_item2_r48 PRG_RP_ITEM_PX=8192 PRG_RP_RING_MB=48
run a_nodeps_nofence PRG_RP_FLAGS=6
for vt in h:224 b:256 b0:256 i:224 j:480 k:352; do
  v=${vt%%:*}; T=${vt##*:}
  run ${v} PRG_LIB_PATH=$L/libprg_$v.so
  run ${v}_r48 PRG_LIB_PATH=$L/libprg_$v.so PRG_RP_RING_MB=48
  run ${v}_item2_r48 PRG_LIB_PATH=$L/libprg_$v.so PRG_RP_ITEM_PX=$((2*16*${T:-256})) PRG_RP_RING_MB=48
done
