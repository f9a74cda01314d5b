#!/bin/bash
# Round-2 GPU call 1: baseline -m gpu suite, A/B of the merged round-1 experiments, fresh ncu captures
# of the shipped kernels.   gpurun --timeout 1500 -- 'bash tools/gpu_r02_call1.sh'
set -u
mkdir -p gpurun_out
nproc > gpurun_out/c1_nproc.txt; python - <<'PY' >> gpurun_out/c1_nproc.txt
import os
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
try: print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e: print("cpu.max n/a", e)
PY
echo "=== gpu suite (baseline) ==="
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() {
  local name=$1; shift
  echo "=== $name ==="
  if [ "$name" != baseline ]; then env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1; fi
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
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
echo "=== full chain d0.5 baseline / strip2 ==="
for v in NONE STRIP2; do
  env JXLGPU_$v=1 timeout 300 python bench.py --workload 8k-d0.5-full --steps 10 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ab_full_$v.json 2> gpurun_out/ab_full_$v.err
  python - $v <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab_full_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], d["ms_per_step"], d["roofline"]["kernel_ms"])
PY
done
echo "=== ncu ==="
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_8k-d1.csv \
    python tools/profile_run.py 8k-d1 3 f32 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'idct|filter_strip' -s 4 -c 4 -f -o gpurun_out/r02_full_8k-d1 \
    python tools/profile_run.py 8k-d1 2 f32 > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'filter_strip' -s 1 -c 1 -f -o gpurun_out/r02_full_8k-d05 \
    python tools/profile_run.py 8k-d0.5-full 2 f32 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep
