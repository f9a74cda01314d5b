#!/usr/bin/env python3
"""Times frames that end in the stages behind the filters (upsampling, noise) on the device:
render_device with CUDA events on the launch stream, frame_begin (which generates the noise planes) by wall clock.
    python tools/measure_post_stages.py"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import jxl_workload as wl  # noqa: E402
from libjxl_b200 import pipeline  # noqa: E402

NOISE_LUT = (0.001, 0.0068, 0.0039, 0.0049, 0.0059, 0.0078, 0.0088, 0.0107)


def main():
    wts = np.load(ROOT / "tests" / "golden" / "upsampling_weights.npz")
    pipe = pipeline.TransformPipeline(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    base, coeffs = wl.synthetic_frame(3840, 2160, seed=3, strategies="0,1,2,3,4,5", epf_iters=1)
    dev = torch.from_numpy(coeffs).cuda()
    for n, noise in ((1, 0), (1, 1), (2, 0), (2, 1), (4, 0)):
        import copy
        desc = copy.copy(base)
        if n > 1:
            desc.upsampling, desc.upsampling_weights = n, wts[f"weights{n}"]
        if noise:
            desc.noise, desc.noise_lut = 1, NOISE_LUT
        pipe.set_device_coefficients([dev[c].data_ptr() for c in range(3)])
        out = torch.empty((desc.out_ysize, desc.out_xsize, 3), dtype=torch.float32, device="cuda")
        t_fb = []
        for _ in range(6):
            torch.cuda.synchronize()
            a = time.perf_counter()
            pipe.frame_begin(desc)
            pipe.synchronize()
            t_fb.append((time.perf_counter() - a) * 1e3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            pipe.render_device(out.data_ptr(), desc.out_row_bytes, stream.cuda_stream)
        e0.record(stream)
        for _ in range(10):
            pipe.render_device(out.data_ptr(), desc.out_row_bytes, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"coded 3840x2160, upsampling {n}, noise {noise}: output {desc.out_xsize}x{desc.out_ysize}, render {ms:.3f} ms "
              f"= {desc.out_xsize * desc.out_ysize / ms / 1e3:.0f} Mpixel/s of output; frame_begin {np.median(t_fb):.3f} ms", flush=True)
        pipe.set_device_coefficients(None)
    pipe.close()


if __name__ == "__main__":
    main()
