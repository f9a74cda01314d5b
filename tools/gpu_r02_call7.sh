#!/bin/bash
# Round-2 GPU call 7: engaged-block rotation, TMA-staged idct8 (A/B), sliced large IDCT, plan sigma.
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
PY
}
for w in 8k-d1 8k-d0.5-full 4k-d1 4k-all27; do
echo "=== $w (TMA idct8) ==="
timeout 900 python bench.py --workload $w --no-cpu-baseline --no-variants > gpurun_out/c7_$w.json 2> gpurun_out/c7_$w.err; show c7_$w
done
echo "=== 8k-d1, JXLGPU_IDCT8_TMA=0 ==="
JXLGPU_IDCT8_TMA=0 timeout 900 python bench.py --workload 8k-d1 --no-cpu-baseline --no-variants > gpurun_out/c7_8k-d1_notma.json 2> gpurun_out/c7_notma.err; show c7_8k-d1_notma
echo "=== int16 frame: 8k-d1 srgb8? (4k-all27 is int16) ==="
echo "=== ncu full: idct8 tma + filter ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'idct8_tma_kernel|filter_strip_kernel' -s 2 -c 2 -f -o gpurun_out/r02_full_tma_rot_8k-d1 \
    python tools/profile_run.py 8k-d1 2 f32 > gpurun_out/ncu_c7.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'idct_large_kernel|plan_kernel' -c 3 -f -o gpurun_out/r02_full_large_all27 \
    python tools/profile_run.py 4k-all27 1 f32 > gpurun_out/ncu_c7b.log 2>&1
tail -3 gpurun_out/ncu_c7b.log
