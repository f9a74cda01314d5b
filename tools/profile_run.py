#!/usr/bin/env python3
"""Small driver for ncu: prepares one frame and renders it a few times device-resident.
    ncu ... python tools/profile_run.py [workload] [renders]"""
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
    fr, _ = bench.prepare_frame(name, 0, 1, lambda: None)
    desc, coeffs = fr["desc"], fr["coeffs"]
    dev = torch.from_numpy(coeffs).cuda()
    out = torch.empty((desc.ysize, desc.xsize, 3), dtype=torch.float32, device="cuda")
    pipe = pipeline.TransformPipeline(0)
    pipe.set_device_coefficients([dev[c].data_ptr() for c in range(3)])
    pipe.frame_begin(desc)
    for _ in range(n):
        pipe.render_device(out.data_ptr(), desc.xsize * 12, 0)
    pipe.synchronize()
    print("rendered", n, "x", name, float(out.float().mean()))


if __name__ == "__main__":
    main()
