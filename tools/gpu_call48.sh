#!/bin/bash
mkdir -p gpurun_out/c48
O=gpurun_out/c48
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/profile_forward.py --batch 2 --size 256 > $O/memcheck_unet.log 2>&1; echo "unet rc=$?" | tee -a $O/memcheck_unet.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_geometry.py > $O/memcheck_geometry.log 2>&1; echo "geometry rc=$?" | tee -a $O/memcheck_geometry.log
grep -E "ERROR SUMMARY" $O/memcheck_unet.log $O/memcheck_geometry.log
tail -6 $O/memcheck_geometry.log
