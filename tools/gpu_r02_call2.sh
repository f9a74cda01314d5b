#!/bin/bash
# Round-2 GPU call 2: the fused kernel -- gpu suite, A/B against the two-kernel path, ncu.
set -u
mkdir -p gpurun_out
echo "=== gpu suite (fused default) ==="
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
except Exception as e:
    print("  no result:", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:]); sys.exit(0)
print(f"  f32 : {d['ms_per_step']:.3f} ms/step  {d['roofline']['kernel_ms']}  e2e {d['e2e']['value']:.0f} parity {d['parity']}")
v = (d.get("variants") or {}).get("srgb8")
if v: print(f"  u8  : {v['ms_per_step']:.3f} ms/step  {v['kernel_ms']}   e2e {v['e2e']['value']:.0f} parity {v['parity']}")
PY
}
echo "=== bench fused ==="
JXLGPU_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c2_fused.json 2> gpurun_out/c2_fused.err; show c2_fused
grep -m3 "fused chain" gpurun_out/c2_fused.err
echo "=== bench two-kernel ==="
JXLGPU_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/c2_two.json 2> gpurun_out/c2_two.err; show c2_two
echo "=== d0.5 (mask 16 after default gab/epf) and 4k ==="
timeout 600 python bench.py --workload 4k-d1 --steps 20 --warmup 5 --no-cpu-baseline --no-variants > gpurun_out/c2_4k.json 2> gpurun_out/c2_4k.err; show c2_4k
echo "=== ncu ==="
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fused_8k-d1.csv \
    python tools/profile_run.py 8k-d1 3 f32 > gpurun_out/ncu_list2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fused_tile' -s 1 -c 1 -f -o gpurun_out/r02_full_fused_8k-d1 \
    python tools/profile_run.py 8k-d1 2 f32 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep
