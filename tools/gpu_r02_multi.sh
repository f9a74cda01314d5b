#!/bin/bash
# Round-2 multi-GPU call: usage  gpurun --gpus N -- 'bash tools/gpu_r02_multi.sh N [quick]'
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
except Exception as e:
    print("  no result:", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1800:]); sys.exit(0)
print(f"  N={d['n_gpus']} {d['config']['workload'][:36]}: {d['value']:.0f} Mpx/s {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.0f}  parity {str(d['parity'])[:110]}")
print("     ", d['config'].get('parallelism', '')[:150])
v = (d.get("variants") or {}).get("srgb8")
if v: print(f"   u8: {v['value']:.0f} Mpx/s {v['ms_per_step']:.3f} ms/step  e2e {v['e2e']['value']:.0f}")
if "latency_ms" in d: print("     ", d["latency_ms"])
PY
}
run() {  # name, args...
  name=$1; shift
  timeout ${STEP_TIMEOUT:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  show $name
}
echo "=== 8k-d1, gather = copy engines (default) ==="
run mg${N}_8k-d1_ce --workload 8k-d1
if [ "$N" != "8" ]; then
echo "=== 8k-d1, gather = p2p stores fused in the kernel ==="
run mg${N}_8k-d1_p2p --workload 8k-d1 --gather p2p --no-variants
fi
echo "=== 8k-d1, gather = nccl ==="
run mg${N}_8k-d1_nccl --workload 8k-d1 --gather nccl --no-variants
echo "=== 8k-d0.5-full ==="
run mg${N}_8k-d0.5-full --workload 8k-d0.5-full --no-variants
echo "=== 64x1080p replicas ==="
run mg${N}_1080p --workload 64x1080p --steps 10
if [ "$N" = "8" ]; then
echo "=== 16k-d2-epf3 ==="
STEP_TIMEOUT=300 run mg${N}_16k --workload 16k-d2-epf3 --steps 10 --no-variants
fi
