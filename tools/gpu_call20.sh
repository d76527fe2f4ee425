#!/bin/bash
mkdir -p gpurun_out/c20
timeout 600 python -m pytest tests/test_geometry_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/c20/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c20/pytest.log
tail -3 gpurun_out/c20/pytest.log
run() {
  name=$1; shift
  echo "== $name" | tee -a gpurun_out/c20/variants.txt
  env "$@" timeout 100 python tools/bench_geometry.py --only reproject --maps 500 2>&1 | tail -1 | tee -a gpurun_out/c20/variants.txt
}
run default A=1
run ring36 PRG_RP_RING_MB=36
run item4 PRG_RP_ITEM_PX=30720
run item8_r64 PRG_RP_ITEM_PX=61440 PRG_RP_RING_MB=64
run nodeps_nofence PRG_RP_FLAGS=6
timeout 100 python tools/bench_geometry.py --only reproject --maps 512 --h 256 --w 256 2>&1 | tail -1 | tee -a gpurun_out/c20/variants.txt
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused" -c 1 -f -o gpurun_out/c20/rp python tools/profile_geometry.py > gpurun_out/c20/ncu.log 2>&1
