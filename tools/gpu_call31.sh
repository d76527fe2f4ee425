#!/bin/bash
mkdir -p gpurun_out/c31
O=gpurun_out/c31
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_traj_gpu.py -m gpu -q -x -s > $O/pytest_s.log 2>&1; echo "pytest rc=$?" >> $O/pytest_s.log
grep -E "fused vs separate|passed|failed|rc=" $O/pytest_s.log | tail -5
