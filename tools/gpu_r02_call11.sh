#!/bin/bash
# Round-2 GPU call 11: full gpu suite, default bench (NUMA binding) + reference arm, upsampled / progressive integration.
set -u
mkdir -p gpurun_out
echo "=== gpu suite ==="
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines() if l.startswith("{")][-1])
except Exception as e:
    print("  no result:", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-2500:]); sys.exit(0)
km = (d.get('roofline') or {}).get('kernel_ms')
print(f"  {d['config']['workload'][:40]}: {d['value']:.0f} Mpx/s {d['ms_per_step']:.3f} ms/step  {km and {k: round(v,3) for k,v in km.items()}}  e2e {d['e2e']['value']:.0f} parity {d['parity']}")
v = (d.get("variants") or {}).get("srgb8")
if v: print(f"  u8  : {v['ms_per_step']:.3f} ms/step  e2e {v['e2e']['value']:.0f} / other {v.get('e2e_other_submit',{}).get('value')} parity {v['parity']}")
for k in ("cpu_baseline", "t_e2e_decoder", "roofline"):
    if k in d: print("  ", k, json.dumps(d[k])[:1300])
if "variants" in d and d["variants"].get("e2e_other_submit"): print("   e2e other submit", d["variants"]["e2e_other_submit"])
PY
}
echo "=== default bench ==="
timeout 900 python bench.py > gpurun_out/c11_default.json 2> gpurun_out/c11_default.err; show c11_default; grep "bound\|binding" gpurun_out/c11_default.err | head -3
echo "=== reference arm ==="
timeout 600 python bench.py --impl reference > gpurun_out/c11_ref.json 2> gpurun_out/c11_ref.err; tail -c 600 gpurun_out/c11_ref.json
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
