#!/bin/bash
# usage: bash tools/gpu_r02_mg_sm.sh N  -- 8K d1.0 f32 with the gather variants (sm vs ce vs nccl), NUMA-bound ranks
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --no-cpu-baseline --workload 8k-d1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - $name <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines() if l.startswith("{")][-1])
    v = (d.get("variants") or {}).get("srgb8")
    print(f"  {sys.argv[1]}: {d['ms_per_step']:.3f} ms/step {d['value']:.0f} Mpx/s e2e {d['e2e']['value']:.0f}", v and f"u8 {v['ms_per_step']:.3f} ms {v['value']:.0f} e2e {v['e2e']['value']:.0f}", str(d['parity'])[:60])
except Exception as e:
    print("  no result", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
}
run mgsm${N}_sm --gather sm
run mgsm${N}_ce --gather ce --no-variants
run mgsm${N}_nccl --gather nccl --no-variants
grep -h "bound\|binding" gpurun_out/mgsm${N}_sm.err | head -8
