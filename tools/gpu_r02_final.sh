#!/bin/bash
# Round-2 final sanity: full gpu suite, default bench, post-stage timings, smoke.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-variants > gpurun_out/final_default.json 2> gpurun_out/final_default.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/final_default.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f"  f32 {d['ms_per_step']:.3f} ms {d['value']:.0f} Mpx/s e2e {d['e2e']['value']:.0f} launches {d['gpu_launches']} {d['roofline']['kernel_ms']} parity {d['parity']}")
except Exception as e:
    print("no result", e); print(open("gpurun_out/final_default.err").read()[-2000:])
PY
timeout 200 python tools/measure_post_stages.py 2>&1 | tail -5
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
