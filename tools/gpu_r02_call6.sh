#!/bin/bash
# Round-2 GPU call 6: EPF block permutation in the strip kernel (A/B vs call 5), replicas fix.
set -u
mkdir -p gpurun_out
echo "=== gpu parity suite ==="
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
except Exception as e:
    print("  no result:", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-2500:]); sys.exit(0)
km = (d.get('roofline') or {}).get('kernel_ms')
print(f"  {d['config']['workload'][:40]}: {d['value']:.0f} Mpx/s {d['ms_per_step']:.3f} ms/step  {km and {k: round(v,3) for k,v in km.items()}}  e2e {d['e2e']['value']:.0f} parity {d['parity']}")
v = (d.get("variants") or {}).get("srgb8")
if v: print(f"  u8  : {v['ms_per_step']:.3f} ms/step  {v['kernel_ms']}   e2e {v['e2e']['value']:.0f} / other {v.get('e2e_other_submit',{}).get('value')} parity {v['parity']}")
for k in ("latency_ms",):
    if k in d: print("  ", k, json.dumps(d[k])[:600])
PY
}
for w in 8k-d1 8k-d0.5-full 4k-d1 4k-all27; do
echo "=== $w ==="
timeout 900 python bench.py --workload $w --no-cpu-baseline > gpurun_out/c6_$w.json 2> gpurun_out/c6_$w.err; show c6_$w
done
echo "=== 64x1080p replicas N=1 ==="
timeout 900 python bench.py --workload 64x1080p --steps 10 --no-cpu-baseline > gpurun_out/c6_1080p.json 2> gpurun_out/c6_1080p.err; show c6_1080p
echo "=== ncu full: filter 21 ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'filter_strip_kernel' -s 1 -c 1 -f -o gpurun_out/r02_full_filter21_compact \
    python tools/profile_run.py 8k-d1 2 f32 > gpurun_out/ncu_f21.log 2>&1
echo "=== ncu full: filter 31 ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'filter_strip_kernel' -s 1 -c 1 -f -o gpurun_out/r02_full_filter31_compact \
    python tools/profile_run.py 8k-d0.5-full 2 f32 > gpurun_out/ncu_f31.log 2>&1
ls gpurun_out | head -40
