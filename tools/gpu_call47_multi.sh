#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out/c47
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@"; }
run --workload dataset --steps 1 --warmup 1 --pairs 64 --no-cpu-baseline > gpurun_out/c47/bench_dataset_b4_n$N.json 2> gpurun_out/c47/bench_dataset_b4_n$N.err
python - $N <<'PY'
import json,sys
d=json.loads(open("gpurun_out/c47/bench_dataset_b4_n%s.json"%sys.argv[1]).read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['config']['batch_size'], d['config']['device_batch'])
PY
