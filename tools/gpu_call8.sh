#!/bin/bash
mkdir -p gpurun_out/c8
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c8/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8/pytest.log
timeout 600 python -m pytest tests/test_traj_gpu.py tests/test_round2_gpu.py -m gpu -q -s > gpurun_out/c8/pytest_prints.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c8/smoke.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/c8/launches_unet_b32.csv python tools/profile_forward.py --batch 32 > gpurun_out/c8/ncu_launches.log 2>&1
tail -3 gpurun_out/c8/pytest.log; grep -h "teacher-forced\|free-running\|philox\|maskunet\|keep-mask" gpurun_out/c8/pytest_prints.log | head -60; tail -2 gpurun_out/c8/smoke.log
