#!/bin/bash
# Round-2 GPU call 8: TMA idct8 v2 (dequant matrices in shared memory, records two items ahead, conflict-free staged reads) A/B.
set -u
mkdir -p gpurun_out
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
echo "=== gpu tests (parity file) ==="
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for w in 8k-d1 4k-all27 8k-d0.5-full; do
echo "=== $w TMA ==="
timeout 900 python bench.py --workload $w --no-cpu-baseline --no-variants > gpurun_out/c8_$w.json 2> gpurun_out/c8_$w.err; show c8_$w
echo "=== $w JXLGPU_IDCT8_TMA=0 ==="
JXLGPU_IDCT8_TMA=0 timeout 900 python bench.py --workload $w --no-cpu-baseline --no-variants > gpurun_out/c8_${w}_notma.json 2> gpurun_out/c8_${w}_notma.err; show c8_${w}_notma
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'idct8_tma_kernel' -s 1 -c 1 -f -o gpurun_out/r02_full_idct8_tma_v2 \
    python tools/profile_run.py 8k-d1 2 f32 > gpurun_out/ncu_c8.log 2>&1
tail -2 gpurun_out/ncu_c8.log
