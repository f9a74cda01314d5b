#!/bin/bash
# Round-2 GPU call 9: upsampling on the device, integration incl. upsampled frames, idct8 TMA v3 A/B.
set -u
mkdir -p gpurun_out
echo "=== gpu tests: parity + integration ==="
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_integration_libjxl.py -m gpu -x -q 2>&1 | tail -4
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
except Exception as e:
    print("  no result:", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-2500:]); sys.exit(0)
km = (d.get('roofline') or {}).get('kernel_ms')
print(f"  {d['config']['workload'][:40]}: {d['value']:.0f} Mpx/s {d['ms_per_step']:.3f} ms/step  {km and {k: round(v,3) for k,v in km.items()}}  e2e {d['e2e']['value']:.0f} parity {d['parity']}")
PY
}
for w in 8k-d1 4k-d1; do
echo "=== $w JXLGPU_IDCT8_TMA=1 ==="
JXLGPU_IDCT8_TMA=1 timeout 900 python bench.py --workload $w --no-cpu-baseline --no-variants > gpurun_out/c9_$w.json 2> gpurun_out/c9_$w.err; show c9_$w
echo "=== $w JXLGPU_IDCT8_TMA=0 ==="
JXLGPU_IDCT8_TMA=0 timeout 900 python bench.py --workload $w --no-cpu-baseline --no-variants > gpurun_out/c9_${w}_notma.json 2> gpurun_out/c9_${w}_notma.err; show c9_${w}_notma
done
echo "=== fused kernel (staggered scratch) ==="
JXLGPU_FUSED=1 timeout 600 python bench.py --no-cpu-baseline --no-variants > gpurun_out/c9_fused.json 2> gpurun_out/c9_fused.err; show c9_fused
JXLGPU_IDCT8_TMA=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'idct8_tma_kernel' -s 1 -c 1 -f -o gpurun_out/r02_full_idct8_tma_v3 \
    python tools/profile_run.py 8k-d1 2 f32 > gpurun_out/ncu_c9.log 2>&1
tail -2 gpurun_out/ncu_c9.log
