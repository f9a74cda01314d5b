"""Shared helpers for the test-suite."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from libjxl_b200 import abi

GOLDEN = Path(__file__).resolve().parent / "golden"


class _Info:
    pass


class GoldenDump:
    """tests/golden/frame_small.npz viewed like an oracle.ref.FrameDump."""

    def __init__(self):
        z = np.load(GOLDEN / "frame_small.npz")
        self.z = z
        self.info = _Info()
        for k in z.files:
            if k.startswith("info_"):
                v = z[k]
                setattr(self.info, k[5:], v.tolist() if v.ndim else v.item())
        self.ac_strategy, self.raw_quant, self.sharpness = z["ac_strategy"], z["raw_quant"], z["sharpness"]
        self.ytox, self.ytob, self.dc = z["ytox"], z["ytob"], z["dc"]
        self.dequant, self.dequant_offsets, self.coeffs = z["dequant"], z["dequant_offsets"], z["coeffs"]
        self.taps = {k[4:]: z[k] for k in z.files if k.startswith("tap_")}
        self.decoded_default = z["decoded_default"]
        self.sigma_interior = z["sigma_interior"]


def golden_desc(**overrides) -> tuple[abi.FrameDesc, np.ndarray, GoldenDump]:
    from oracle import cpu as ocpu
    g = GoldenDump()
    return ocpu.desc_from_dump(g, **overrides), g.coeffs, g


def ulp_diff(a: np.ndarray, b: np.ndarray) -> int:
    """max distance in float32 ULPs (ordered-integer representation)."""
    ai = a.astype(np.float32).view(np.int32).astype(np.int64)
    bi = b.astype(np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-2147483648) - ai, ai)
    bi = np.where(bi < 0, np.int64(-2147483648) - bi, bi)
    return int(np.abs(ai - bi).max())


TAP_MASKS = {"idct": 0, "gab_epf012": 15, "full": 31}


DC_STAGE_CASES = [(37, 21), (64, 48), (3, 3), (2, 9)]
DC_FACTORS = (3.1 / 4096, 1.7 / 512, 0.9 / 256)
DC_CFL = (0.0117, 0.0, 0.935)


def dc_stage_input(xs: int, ys: int) -> np.ndarray:
    """Seeded quantised DC planes (X, Y, B): smooth gradients (adaptive smoothing engages) with one busy
    quadrant (it must switch itself off there)."""
    rng = np.random.default_rng(1000 * xs + ys)
    yy, xx = np.mgrid[0:ys, 0:xs]
    q = np.stack([np.round(20 * np.sin(xx / 17 + c) + 15 * np.cos(yy / 23 + c) + rng.random((ys, xs)) * 1.2)
                  for c in range(3)]).astype(np.int32)
    q[:, ys // 2:, xs // 2:] += rng.integers(-40, 40, (3, ys - ys // 2, xs - xs // 2))
    return q
