#!/bin/bash
# Round-2 GPU call 13: full gpu suite (noise, upsampling, progressive, gather) + a short bench sanity run.
set -u
mkdir -p gpurun_out
echo "=== gpu suite ==="
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "=== default bench (no cpu baseline) ==="
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c13_default.json 2> gpurun_out/c13_default.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c13_default.json").read().strip().splitlines() if l.startswith("{")][-1])
    v = d["variants"]["srgb8"]
    print(f"  f32 {d['ms_per_step']:.3f} ms {d['value']:.0f} Mpx/s e2e {d['e2e']['value']:.0f} | u8 {v['ms_per_step']:.3f} ms e2e {v['e2e']['value']:.0f} | launches {d['gpu_launches']} parity {d['parity']}")
    print("  t_e2e", json.dumps(d.get("t_e2e_decoder"))[:300])
except Exception as e:
    print("no result", e); print(open("gpurun_out/c13_default.err").read()[-2000:])
PY
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
