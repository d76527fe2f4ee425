#!/bin/bash
mkdir -p gpurun_out/c49
O=gpurun_out/c49
timeout 300 compute-sanitizer --tool racecheck --kernel-regex kns=k_res1x1_gn --error-exitcode 3 python tools/profile_forward.py --batch 3 --size 256 > $O/racecheck_resgn.log 2>&1; echo "racecheck rc=$?" | tee -a $O/racecheck_resgn.log
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard" $O/racecheck_resgn.log | head -5
timeout 200 compute-sanitizer --tool racecheck --kernel-regex kns=k_net_tail --error-exitcode 3 python tools/profile_forward.py --batch 3 --size 128 --net mask > $O/racecheck_tail.log 2>&1; echo "racecheck tail rc=$?" | tee -a $O/racecheck_tail.log
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard" $O/racecheck_tail.log | head -5
