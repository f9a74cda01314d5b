#!/bin/bash
# Round-2 GPU call 3: fused kernel v2 (hoisted step state), integration test on the GPU, full default bench.
set -u
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
except Exception as e:
    print("  no result:", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-2500:]); sys.exit(0)
print(f"  f32 : {d['ms_per_step']:.3f} ms/step  {d['roofline']['kernel_ms']}  e2e {d['e2e']['value']:.0f} parity {d['parity']}")
v = (d.get("variants") or {}).get("srgb8")
if v: print(f"  u8  : {v['ms_per_step']:.3f} ms/step  {v['kernel_ms']}   e2e {v['e2e']['value']:.0f} parity {v['parity']}")
for k in ("cpu_baseline", "t_e2e_decoder"):
    if k in d: print("  ", k, json.dumps(d[k])[:900])
PY
}
echo "=== bench fused v2 ==="
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/c3_fused.json 2> gpurun_out/c3_fused.err; show c3_fused
echo "=== integration test on the GPU ==="
timeout 900 python -m pytest tests/test_integration_libjxl.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
echo "=== ncu fused v2 ==="
ncu --set full --clock-control none --import-source on -k regex:'fused_tile' -s 1 -c 1 -f -o gpurun_out/r02_full_fused_v2_8k-d1 \
    python tools/profile_run.py 8k-d1 2 f32 > gpurun_out/ncu_full4.log 2>&1
echo "=== full default bench (two-kernel path forced: JXLGPU_FUSED=0) with CPU baseline + T_e2e ==="
JXLGPU_FUSED=0 timeout 900 python bench.py > gpurun_out/c3_default.json 2> gpurun_out/c3_default.err; show c3_default
tail -5 gpurun_out/c3_default.err
