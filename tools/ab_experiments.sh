#!/bin/bash
# A/B of the round-1 kernel experiments on ONE gpurun call (branch exp/all):
#   gpurun --timeout 900 -- 'bash tools/ab_experiments.sh'
# Prints per-kernel CUDA-event times of `bench.py` for the baseline and for each switch; correctness of
# every variant is checked first with the -m gpu suite under the same switch.
set -u
mkdir -p gpurun_out
run() {  # name, env assignments...
  local name=$1; shift
  echo "=== $name ==="
  env "$@" python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
k = d["roofline"]["kernel_ms"]
v = d["variants"]["srgb8"]
print(f"  f32 : {d['ms_per_step']:.3f} ms/step  idct8 {k['idct8']:.3f}  filter {k['filter']:.3f}   e2e {d['e2e']['value']:.0f}")
print(f"  u8  : {v['ms_per_step']:.3f} ms/step  idct8 {v['kernel_ms']['idct8']:.3f}  filter {v['kernel_ms']['filter']:.3f}   e2e {v['e2e']['value']:.0f}")
PY
}
run baseline JXLGPU_NONE=1
run strip2 JXLGPU_STRIP2=1
run idct8pipe JXLGPU_IDCT8_PIPE=1
run srgb8special JXLGPU_SRGB8_SPECIAL=1
run all JXLGPU_STRIP2=1 JXLGPU_IDCT8_PIPE=1 JXLGPU_SRGB8_SPECIAL=1
