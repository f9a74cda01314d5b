#!/bin/bash
# Round-2 GPU call 14: ncu launch list of the final default path (shares of the step), timing of an upsampled + noisy frame.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final_8k-d1.csv \
    python tools/profile_run.py 8k-d1 3 f32 > gpurun_out/ncu_list_final.log 2>&1
tail -2 gpurun_out/ncu_list_final.log
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import jxl_workload as wl
from libjxl_b200 import abi, pipeline
wts = np.load("tests/golden/upsampling_weights.npz")
pipe = pipeline.TransformPipeline(0)
for n, w, h, noise in ((2, 3840, 2160, 0), (2, 3840, 2160, 1), (1, 3840, 2160, 1), (1, 3840, 2160, 0)):
    desc, coeffs = wl.synthetic_frame(w, h, seed=3, strategies="0,1,2,3,4,5", epf_iters=1)
    if n > 1:
        desc.upsampling, desc.upsampling_weights = n, wts[f"weights{n}"]
    if noise:
        desc.noise, desc.noise_lut = 1, (0.001, 0.0068, 0.0039, 0.0049, 0.0059, 0.0078, 0.0088, 0.0107)
    dev = torch.from_numpy(coeffs).cuda()
    pipe.set_device_coefficients([dev[c].data_ptr() for c in range(3)])
    out = torch.empty((desc.out_ysize, desc.out_xsize, 3), dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream()
    t_fb = []
    for rep in range(6):
        torch.cuda.synchronize(); a = time.perf_counter()
        pipe.frame_begin(desc); pipe.synchronize(); t_fb.append((time.perf_counter() - a) * 1e3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        pipe.render_device(out.data_ptr(), desc.out_row_bytes, s.cuda_stream)
    e0.record(s)
    for _ in range(10):
        pipe.render_device(out.data_ptr(), desc.out_row_bytes, s.cuda_stream)
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"coded {w}x{h} upsampling {n} noise {noise}: render {ms:.3f} ms -> {desc.out_xsize * desc.out_ysize / ms / 1e3:.0f} Mpx/s of output; frame_begin (incl. noise generation) {np.median(t_fb):.3f} ms")
    pipe.set_device_coefficients(None)
PY
