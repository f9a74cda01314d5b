#!/bin/bash
# Round-2 GPU call 5 (re-entry): everything built so far, measured in one go.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python - <<'PY'
import os
print("affinity cores", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count(), open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "no cpu.max")
PY
echo "=== gpu suite ==="
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
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
for k in ("cpu_baseline", "t_e2e_decoder", "latency_ms", "roofline"):
    if k in d: print("  ", k, json.dumps(d[k])[:1500])
if "variants" in d and d["variants"].get("e2e_other_submit"): print("   e2e other submit", d["variants"]["e2e_other_submit"])
PY
}
echo "=== default bench ==="
timeout 900 python bench.py > gpurun_out/c5_default.json 2> gpurun_out/c5_default.err; show c5_default
echo "=== reference arm ==="
timeout 600 python bench.py --impl reference > gpurun_out/c5_ref.json 2> gpurun_out/c5_ref.err; tail -c 1500 gpurun_out/c5_ref.json
echo "=== fused (opt-in) ==="
JXLGPU_FUSED=1 timeout 600 python bench.py --no-cpu-baseline --no-variants > gpurun_out/c5_fused.json 2> gpurun_out/c5_fused.err; show c5_fused
echo "=== 64x1080p replicas N=1 ==="
timeout 900 python bench.py --workload 64x1080p --steps 10 --no-cpu-baseline > gpurun_out/c5_1080p.json 2> gpurun_out/c5_1080p.err; show c5_1080p
echo "=== 4k-all27 ==="
timeout 900 python bench.py --workload 4k-all27 --no-cpu-baseline --no-variants > gpurun_out/c5_all27.json 2> gpurun_out/c5_all27.err; show c5_all27
echo "=== 8k-d0.5-full ==="
timeout 900 python bench.py --workload 8k-d0.5-full --no-cpu-baseline --no-variants > gpurun_out/c5_d05.json 2> gpurun_out/c5_d05.err; show c5_d05
echo "=== 4k-d1 ==="
timeout 900 python bench.py --workload 4k-d1 --no-cpu-baseline --no-variants > gpurun_out/c5_4k.json 2> gpurun_out/c5_4k.err; show c5_4k
echo "=== ncu launch list (two-kernel) ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_8k-d1.csv \
    python tools/profile_run.py 8k-d1 3 f32 > gpurun_out/ncu_list.log 2>&1
echo "=== ncu full: two-kernel path ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'idct8_kernel|filter_strip_kernel|idct_mid' -s 3 -c 3 -f -o gpurun_out/r02_full_two_8k-d1 \
    python tools/profile_run.py 8k-d1 2 f32 > gpurun_out/ncu_full_two.log 2>&1
echo "=== ncu full: fused ==="
JXLGPU_FUSED=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fused' -s 1 -c 1 -f -o gpurun_out/r02_full_fused_8k-d1 \
    python tools/profile_run.py 8k-d1 2 f32 > gpurun_out/ncu_full_fused.log 2>&1
ls -la gpurun_out/
