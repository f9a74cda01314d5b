#!/bin/bash
# Round-2 GPU call 12 (2 GPUs): the auto-gather code path with the sm threshold lowered to 2 ranks, DC stage timing.
set -u
mkdir -p gpurun_out
echo "=== DC stage on the device (rank 0 only) ==="
timeout 300 python tools/measure_dc_stage.py 8k-d1 2>&1 | tail -2
echo "=== auto gather, sm forced from 2 ranks (code path of N>=4) ==="
BENCH_AUTO_SM_MIN_WORLD=2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --no-cpu-baseline > gpurun_out/c12_auto_sm.json 2> gpurun_out/c12_auto_sm.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c12_auto_sm.json").read().strip().splitlines() if l.startswith("{")][-1])
    v = d["variants"]["srgb8"]
    print(f"  f32 {d['ms_per_step']:.3f} ms  {d['config']['parallelism'][:120]}")
    print(f"  u8  {v['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.0f} / {v['e2e']['value']:.0f}  parity {d['parity']} {v['parity']}")
except Exception as e:
    print("no result", e); print(open("gpurun_out/c12_auto_sm.err").read()[-2000:])
PY
echo "=== plain auto at 2 ranks ==="
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --no-cpu-baseline --no-variants > gpurun_out/c12_auto.json 2> gpurun_out/c12_auto.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c12_auto.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"  f32 {d['ms_per_step']:.3f} ms  {d['config']['parallelism'][:120]}  e2e {d['e2e']['value']:.0f}")
except Exception as e:
    print("no result", e); print(open("gpurun_out/c12_auto.err").read()[-2000:])
PY
