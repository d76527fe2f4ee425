"""profiles/conv_traffic.json from an ncu launch list of one U-Net evaluation.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv \
        --log-file gpurun_out/launches_unet_b32.csv python tools/profile_forward.py --batch 32
    python tools/conv_traffic.py gpurun_out/launches_unet_b32.csv profiles/r2_launches_unet_b32.csv

Sums dram__bytes_read.sum + dram__bytes_write.sum over the k_conv2 launches (bench.py's `roofline.traffic`).
"""
import csv
import json
import os
import shutil
import sys

src = sys.argv[1]
dst = sys.argv[2] if len(sys.argv) > 2 else None
rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
per = {}
for r in rows:
    k = int(r[0])
    per.setdefault(k, {"name": r[4]})[r[12]] = float(r[14].replace(",", ""))
conv = [v for v in per.values() if "k_conv2" in v["name"]]
total = sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in conv)
allb = sum(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0) for v in per.values())
t_all = sum(v.get("gpu__time_duration.sum", 0) for v in per.values())
t_conv = sum(v.get("gpu__time_duration.sum", 0) for v in conv)
print("launches %d (conv %d); DRAM bytes conv %.3f GB, all %.3f GB; time conv %.3f ms of %.3f ms (cold, serialised)"
      % (len(per), len(conv), total / 1e9, allb / 1e9, t_conv / 1e6, t_all / 1e6))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if dst:
    shutil.copyfile(src, dst)
    out = {"batch": 32, "size": 256, "dram_bytes_per_unet_eval": int(total),
           "dram_bytes_per_unet_eval_all_kernels": int(allb), "conv_launches": len(conv), "launches": len(per),
           "source": "%s: sum of dram__bytes_read.sum + dram__bytes_write.sum over the %d k_conv2 launches of one U-Net "
                     "evaluation (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                     "--clock-control none, tools/profile_forward.py --batch 32); earlier in round 2: 13 699 601 408, "
                     "round 1: 15 032 200 000" % (os.path.relpath(dst, root), len(conv))}
    with open(os.path.join(root, "profiles", "conv_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
