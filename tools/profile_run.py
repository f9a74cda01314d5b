#!/usr/bin/env python3
"""Small driver for ncu: prepares one frame and renders it a few times device-resident.
    ncu ... python tools/profile_run.py [workload] [renders] [f32|srgb8] [sparse]
(`sparse`: one extra host-fed frame through jxlgpu_submit_groups_sparse, for the scatter kernel)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
from libjxl_b200 import abi, pipeline  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "8k-d1"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    kind = sys.argv[3] if len(sys.argv) > 3 else "f32"
    fr, _ = bench.prepare_frame(name, 0, 1, lambda: None)
    desc, coeffs = fr["desc"], fr["coeffs"]
    if kind == "srgb8":
        desc.out_format, desc.stage_mask = abi.OUT_RGB_U8, abi.STAGE_SRGB
    dev = torch.from_numpy(coeffs).cuda()
    out = torch.empty((desc.ysize, desc.xsize, 3), dtype=torch.float32 if kind == "f32" else torch.uint8, device="cuda")
    pipe = pipeline.TransformPipeline(0)
    pipe.set_device_coefficients([dev[c].data_ptr() for c in range(3)])
    pipe.frame_begin(desc)
    for _ in range(n):
        pipe.render_device(out.data_ptr(), desc.out_row_bytes, 0)
    pipe.synchronize()
    if "sparse" in sys.argv[4:]:
        pipe.decode_frame(desc, coeffs, sparse=True)
    print("rendered", n, "x", name, float(out.float().mean()))


if __name__ == "__main__":
    main()
