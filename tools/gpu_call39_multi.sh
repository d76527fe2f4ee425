#!/bin/bash
# multi-GPU evidence for the three workloads (launched like the driver does: torchrun, one rank per GPU)
N=${1:-2}
mkdir -p gpurun_out/c39
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@"; }
run --workload geometry --steps 2 --warmup 3 > gpurun_out/c39/bench_geometry_n$N.json 2> gpurun_out/c39/bench_geometry_n$N.err
run --workload dataset --steps 1 --warmup 1 --pairs 32 > gpurun_out/c39/bench_dataset_b4_n$N.json 2> gpurun_out/c39/bench_dataset_b4_n$N.err
run --workload dataset --steps 1 --warmup 1 --pairs 64 --batch 32 > gpurun_out/c39/bench_dataset_b32_n$N.json 2> gpurun_out/c39/bench_dataset_b32_n$N.err
run --steps 1 --warmup 3 --no-e2e > gpurun_out/c39/bench_pairs_n$N.json 2> gpurun_out/c39/bench_pairs_n$N.err
for f in gpurun_out/c39/*_n$N.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['n_gpus'], d['value'], d['unit'])
except Exception as e: print(sys.argv[1],'ERR',e)
PY
done
