#!/usr/bin/env python3
"""Times the optional DC stage on the device (dc_dequant_kernel + dc_smooth_kernel inside
jxlgpu_frame_begin) against handing over finished float DC planes, on the 8K bench frame.
    python tools/measure_dc_stage.py [workload]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from libjxl_b200 import pipeline  # noqa: E402
from tests import support  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "8k-d1"
    fr, _ = bench.prepare_frame(name, 0, 1, lambda: None)
    desc, coeffs = fr["desc"], fr["coeffs"]
    dev = torch.from_numpy(coeffs).cuda()
    pipe = pipeline.TransformPipeline(0)
    pipe.set_device_coefficients([dev[c].data_ptr() for c in range(3)])
    # quantised DC + per-DC-group multipliers, and what the host would have made of them (oracle: DequantDC per
    # 2048x2048-px DC group, then AdaptiveDCSmoothing over the whole image)
    from oracle import cpu
    yb, xb = desc.ysize_blocks, desc.xsize_blocks
    q = support.dc_stage_input(xb, yb)
    gm = (0.5 ** (np.arange(((yb + 255) // 256) * ((xb + 255) // 256)) % 4)).astype(np.float32).reshape((yb + 255) // 256, -1)
    dc = np.zeros((3, yb, xb), np.float32)
    for gy in range(gm.shape[0]):
        for gx in range(gm.shape[1]):
            sl = (slice(None), slice(gy * 256, (gy + 1) * 256), slice(gx * 256, (gx + 1) * 256))
            dc[sl] = cpu.dequant_dc(q[sl], support.DC_FACTORS, float(gm[gy, gx]), support.DC_CFL)
    dc = cpu.adaptive_dc_smoothing(dc, support.DC_FACTORS)
    out = torch.empty((desc.ysize, desc.xsize, 3), dtype=torch.float32, device="cuda")

    def run(n):
        t0 = time.perf_counter()
        for _ in range(n):
            pipe.frame_begin(desc)
            pipe.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    desc.dc = dc
    pipeline.pin_side_info(desc)
    run(3)
    host_ms = run(20)
    pipe.render_device(out.data_ptr(), desc.out_row_bytes, 0)
    pipe.synchronize()
    want = out.clone()
    desc.quant_dc, desc.dc_group_mul, desc.dc_smoothing = q, gm, 1
    desc.dc_factors, desc.dc_cfl_factors = support.DC_FACTORS, support.DC_CFL
    desc.dc = None
    pipeline.pin_side_info(desc)
    run(3)
    dev_ms = run(20)
    pipe.render_device(out.data_ptr(), desc.out_row_bytes, 0)
    pipe.synchronize()
    print(f"frame_begin + sync, {name}: float DC planes from the host {host_ms:.3f} ms, quantised DC + device DC stage "
          f"{dev_ms:.3f} ms (dequantisation + adaptive smoothing of {desc.xsize_blocks * desc.ysize_blocks} blocks on the GPU); "
          f"pixels identical: {bool(torch.equal(out, want))}")


if __name__ == "__main__":
    main()
