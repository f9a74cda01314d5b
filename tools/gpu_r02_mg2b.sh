#!/bin/bash
N=2
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --no-cpu-baseline --no-variants "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  grep "e2e phases" gpurun_out/$name.err | head -4; python -c "
import json;d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]);print('$name', round(d['value']), 'e2e', round(d['e2e']['value']))"; }
echo "=== drop symm before e2e ==="; run mg2b_drop --workload 8k-d1
echo "=== keep symm ==="; BENCH_KEEP_SYMM=1 run mg2b_keep --workload 8k-d1
