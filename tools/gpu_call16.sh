#!/bin/bash
mkdir -p gpurun_out/c16
timeout 600 python -m pytest tests/test_geometry_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/c16/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c16/pytest.log
tail -3 gpurun_out/c16/pytest.log
run() {
  name=$1; shift
  echo "== $name" | tee -a gpurun_out/c16/variants.txt
  env "$@" timeout 100 python tools/bench_geometry.py --only reproject --maps 500 2>&1 | tail -1 | tee -a gpurun_out/c16/variants.txt
}
L=$PWD/pointreggpt_b200
run a_256x3 A=1
run a_256x3_item8k PRG_RP_ITEM_PX=8192
run a_256x3_item8k_r48 PRG_RP_ITEM_PX=8192 PRG_RP_RING_MB=48
run a_nodeps_nofence PRG_RP_FLAGS=6
run b_256x2_pf PRG_LIB_PATH=$L/libprg_b.so
run b_256x2_pf_item8k_r48 PRG_LIB_PATH=$L/libprg_b.so PRG_RP_ITEM_PX=8192 PRG_RP_RING_MB=48
run c_384x2 PRG_LIB_PATH=$L/libprg_c.so
run c_384x2_r48 PRG_LIB_PATH=$L/libprg_c.so PRG_RP_RING_MB=48
run d_768x1 PRG_LIB_PATH=$L/libprg_d.so
run d_768x1_item24k_r48 PRG_LIB_PATH=$L/libprg_d.so PRG_RP_ITEM_PX=24576 PRG_RP_RING_MB=48
run d_768x1_nodeps_nofence PRG_LIB_PATH=$L/libprg_d.so PRG_RP_FLAGS=6
run e_512x1_pf PRG_LIB_PATH=$L/libprg_e.so
run e_512x1_pf_item16k PRG_LIB_PATH=$L/libprg_e.so PRG_RP_ITEM_PX=16384
