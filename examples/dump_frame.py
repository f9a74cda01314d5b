#!/usr/bin/env python3
"""Write the binary frame dump examples/host_feed.cc reads: the jxlgpu_frame struct (pointer fields are
re-pointed by the reader), the side-information planes, then the coefficient groups.

    python examples/dump_frame.py frame.bin [width height] [f32|srgb8]

The frame is the seeded synthetic all-strategy frame of jxl_workload.synthetic_frame."""
from __future__ import annotations

import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import jxl_workload as wl  # noqa: E402
from libjxl_b200 import abi  # noqa: E402


def write_dump(path, desc: abi.FrameDesc, coeffs: np.ndarray) -> None:
    yb, xb = desc.ysize_blocks, desc.xsize_blocks

    def arr(f, a, dtype):
        a = np.ascontiguousarray(a, dtype).ravel()
        f.write(struct.pack("<Q", a.size))
        f.write(a.tobytes())

    with open(path, "wb") as f:
        f.write(bytes(desc.to_struct()))
        arr(f, desc.ac_strategy, np.uint8)
        arr(f, desc.raw_quant, np.int32)
        arr(f, desc.epf_sharpness if desc.epf_sharpness is not None else np.zeros(0), np.uint8)
        arr(f, desc.ytox, np.int8)
        arr(f, desc.ytob, np.int8)
        arr(f, desc.dc, np.float32)
        arr(f, desc.dequant, np.float32)
        want = np.int16 if desc.ac_type == abi.AC_INT16 else np.int32
        for g in range(desc.num_groups):
            n = desc.group_ncoeff(g)
            for c in range(3):
                f.write(np.ascontiguousarray(coeffs[c, g, :n], want).tobytes())
    assert yb * xb == desc.ac_strategy.size


def main() -> int:
    out = sys.argv[1]
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1100, 777)
    kind = sys.argv[4] if len(sys.argv) > 4 else "f32"
    desc, coeffs = wl.synthetic_frame(w, h, seed=w + h)
    if kind == "srgb8":
        desc.out_format, desc.stage_mask = abi.OUT_RGB_U8, abi.STAGE_SRGB
    write_dump(out, desc, coeffs)
    print(f"wrote {out}: {w}x{h}, {desc.num_groups} groups, out_format {desc.out_format}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
