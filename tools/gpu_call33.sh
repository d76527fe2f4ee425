#!/bin/bash
mkdir -p gpurun_out/c33
O=gpurun_out/c33
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_traj_gpu.py -m gpu -q -s > $O/pytest_s.log 2>&1
grep -E "fused vs separate|rel|free|final|teacher" $O/pytest_s.log | head -40
