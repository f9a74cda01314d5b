#!/bin/bash
# Round-2 GPU call 4: full gpu suite (both device paths), default bench with T_e2e, the other BASELINE configs at N=1.
set -u
mkdir -p gpurun_out
echo "=== gpu suite ==="
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
except Exception as e:
    print("  no result:", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-2500:]); sys.exit(0)
print(f"  {d['config']['workload'][:60]}: {d['value']:.0f} Mpx/s {d['ms_per_step']:.3f} ms/step  {(d.get('roofline') or {}).get('kernel_ms')}  e2e {d['e2e']['value']:.0f} parity {d['parity']}")
v = (d.get("variants") or {}).get("srgb8")
if v: print(f"  u8  : {v['ms_per_step']:.3f} ms/step  {v['kernel_ms']}   e2e {v['e2e']['value']:.0f} / other {v.get('e2e_other_submit',{}).get('value')} parity {v['parity']}")
for k in ("cpu_baseline", "t_e2e_decoder", "latency_ms"):
    if k in d: print("  ", k, json.dumps(d[k])[:1200])
if "variants" in d and d["variants"].get("e2e_other_submit"): print("   e2e other submit", d["variants"]["e2e_other_submit"])
PY
}
echo "=== default bench ==="
timeout 900 python bench.py > gpurun_out/c4_default.json 2> gpurun_out/c4_default.err; show c4_default
echo "=== reference arm ==="
timeout 600 python bench.py --impl reference > gpurun_out/c4_ref.json 2> gpurun_out/c4_ref.err; tail -c 1500 gpurun_out/c4_ref.json
echo "=== 64x1080p replicas N=1 ==="
timeout 900 python bench.py --workload 64x1080p --steps 10 > gpurun_out/c4_1080p.json 2> gpurun_out/c4_1080p.err; show c4_1080p
echo "=== 4k-all27 ==="
timeout 900 python bench.py --workload 4k-all27 --no-cpu-baseline --no-variants > gpurun_out/c4_all27.json 2> gpurun_out/c4_all27.err; show c4_all27
echo "=== 8k-d0.5-full ==="
timeout 900 python bench.py --workload 8k-d0.5-full --no-cpu-baseline --no-variants > gpurun_out/c4_d05.json 2> gpurun_out/c4_d05.err; show c4_d05
echo "=== 4k-d1 ==="
timeout 900 python bench.py --workload 4k-d1 --no-cpu-baseline --no-variants > gpurun_out/c4_4k.json 2> gpurun_out/c4_4k.err; show c4_4k
