#!/bin/bash
mkdir -p gpurun_out/c19
timeout 600 python -m pytest tests/test_geometry_gpu.py tests/test_golden_gpu.py tests/test_round2_gpu.py -m gpu -q -x > gpurun_out/c19/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c19/pytest.log
tail -3 gpurun_out/c19/pytest.log
run() {
  name=$1; shift
  echo "== $name" | tee -a gpurun_out/c19/variants.txt
  env "$@" timeout 100 python tools/bench_geometry.py --only reproject --maps 500 2>&1 | tail -1 | tee -a gpurun_out/c19/variants.txt
}
L=$PWD/pointreggpt_b200
run default A=1
run ring24 PRG_RP_RING_MB=24
run ring36 PRG_RP_RING_MB=36
run ring64 PRG_RP_RING_MB=64
run item1 PRG_RP_ITEM_PX=7680
run item4 PRG_RP_ITEM_PX=30720
run ahead2 PRG_RP_AHEAD=2
run nodeps_nofence PRG_RP_FLAGS=6
run nofence_variant PRG_LIB_PATH=$L/libprg_nf.so
run small256 A=1
timeout 100 python tools/bench_geometry.py --only reproject --maps 512 --h 256 --w 256 2>&1 | tail -1 | tee -a gpurun_out/c19/variants.txt
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c19/bench_geometry.json 2> gpurun_out/c19/bench_geometry.err
tail -c 1200 gpurun_out/c19/bench_geometry.json
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused" -c 1 -f -o gpurun_out/c19/rp python tools/profile_geometry.py > gpurun_out/c19/ncu.log 2>&1
