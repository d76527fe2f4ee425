#!/bin/bash
mkdir -p gpurun_out/c21
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c21/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c21/pytest.log
tail -3 gpurun_out/c21/pytest.log
run() {
  name=$1; shift
  echo "== $name" | tee -a gpurun_out/c21/variants.txt
  env "$@" timeout 100 python tools/bench_geometry.py --only reproject --maps 500 2>&1 | tail -1 | tee -a gpurun_out/c21/variants.txt
}
run default A=1
run item2 PRG_RP_ITEM_PX=15360
run ring40 PRG_RP_RING_MB=40
run ring56 PRG_RP_RING_MB=56
timeout 100 python tools/bench_geometry.py --only reproject --maps 512 --h 256 --w 256 2>&1 | tail -1 | tee -a gpurun_out/c21/variants.txt
PRG_RP_ITEM_PX=15360 timeout 100 python tools/bench_geometry.py --only reproject --maps 512 --h 256 --w 256 2>&1 | tail -1 | tee -a gpurun_out/c21/variants.txt
timeout 100 python tools/bench_geometry.py --only reproject --maps 32 --h 256 --w 256 2>&1 | tail -1 | tee -a gpurun_out/c21/variants.txt
timeout 300 python bench.py --workload geometry --steps 2 --warmup 3 > gpurun_out/c21/bench_geometry.json 2> gpurun_out/c21/bench_geometry.err
tail -c 700 gpurun_out/c21/bench_geometry.json
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_reproject_fused" -c 1 -f -o gpurun_out/c21/rp python tools/profile_geometry.py > gpurun_out/c21/ncu.log 2>&1
