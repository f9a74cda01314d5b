#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Run where /root/reference exists:

    python tests/golden/make_golden.py

Fixtures
  transforms.npz   jxl::TransformToPixels + LowestFrequenciesFromDC for all 27 AcStrategy types
                   (strict build: -ffp-contract=off, i.e. FMAs exactly where the source says so).
                   Inputs are NOT stored: tests re-create them from the seeds below.
  frame_small.npz  a 264x72 VarDCT d1.0 e7 frame (gab on, epf 3 forced): the reference decoder's
                   coefficient hand-off (quantised coefficients + all side info) and the reference
                   pixels: full decode (default build, public API) and stage taps of the strict
                   build rendered with the reference's own DecodeGroupForRoundtrip + stages.
  outputs_small.npz the same frame through the reference's FromLinearStage (sRGB) and WriteToOutputStage
                   in every packed pixel format (strict build); `--outputs-only` regenerates just this.
  dc_stage.npz     DequantDC + AdaptiveDCSmoothing outputs (the step in front of the path)
                   for seeded inputs; `--dc-only` regenerates just this.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import jxl_workload as wl  # noqa: E402
from libjxl_b200 import abi  # noqa: E402
from oracle import ref  # noqa: E402

HERE = Path(__file__).resolve().parent
TRANSFORM_SEED = 20260922
FRAME = dict(w=264, h=72, distance=1.0, effort=7, gaborish=1, epf=3, seed=99, kind="photo")
TAPS = {"idct": 0, "gab_epf012": 15, "full": 31}


def transform_inputs(strategy: int):
    rng = np.random.default_rng(TRANSFORM_SEED + strategy)
    r, c = abi.COVERED_Y[strategy] * 8, abi.COVERED_X[strategy] * 8
    coeffs = rng.laplace(0, 1.0, r * c).astype(np.float32)
    dc = rng.normal(0, 1, (abi.COVERED_Y[strategy], abi.COVERED_X[strategy])).astype(np.float32)
    return coeffs, dc


OUTPUTS = {"srgb_f32": (abi.OUT_RGB_F32, -33), "srgb_u8": (abi.OUT_RGB_U8, -33), "srgb_rgba8": (abi.OUT_RGBA_U8, -33),
           "srgb_u16": (abi.OUT_RGB_U16, -33), "srgb_f16": (abi.OUT_RGB_F16, -33), "linear_u8": (abi.OUT_RGB_U8, -1),
           "linear_f16": (abi.OUT_RGB_F16, -1)}


def outputs() -> None:
    """outputs_small.npz: the frame of frame_small.npz rendered by the strict reference build through
    its FromLinear (sRGB) and WriteToOutput stages in every packed format."""
    ref.use_variant("strict")
    data = np.load(HERE / "frame_small.npz")["jxl"].tobytes()
    fr = ref.Frame(data, 1)
    fx = {}
    for name, (fmt, mask) in OUTPUTS.items():
        img, _ = fr.render_out(mask, fmt)
        fx[name] = img.view(np.uint16) if img.dtype == np.float16 else img
    fr.close()
    np.savez_compressed(HERE / "outputs_small.npz", **fx)
    print("outputs_small.npz", (HERE / "outputs_small.npz").stat().st_size, "bytes")


def dc_stage() -> None:
    """dc_stage.npz: jxl::DequantDC + jxl::AdaptiveDCSmoothing (strict build) on the seeded inputs of
    tests/support.py:dc_stage_input (inputs are not stored)."""
    sys.path.insert(0, str(ROOT))
    from tests import support
    ref.use_variant("strict")
    fx = {}
    for xs, ys in support.DC_STAGE_CASES:
        q = support.dc_stage_input(xs, ys)
        for mul in (1.0, 0.25):
            fx[f"dequant_{xs}x{ys}_mul{mul}"] = ref.dequant_dc(q, support.DC_FACTORS, mul, support.DC_CFL)
        fx[f"smooth_{xs}x{ys}"] = ref.adaptive_dc_smoothing(fx[f"dequant_{xs}x{ys}_mul1.0"], support.DC_FACTORS, 2)
    np.savez_compressed(HERE / "dc_stage.npz", **fx)
    print("dc_stage.npz", (HERE / "dc_stage.npz").stat().st_size, "bytes")


def main() -> int:
    if "--outputs-only" in sys.argv:
        outputs()
        return 0
    if "--dc-only" in sys.argv:
        dc_stage()
        return 0
    ref.use_variant("strict")
    out = {}
    for s in range(27):
        r, c = abi.COVERED_Y[s] * 8, abi.COVERED_X[s] * 8
        coeffs, dc = transform_inputs(s)
        out[f"px_{s}"] = ref.transform_to_pixels(s, coeffs, r, c)
        llf = ref.llf_from_dc(s, dc, np.zeros(r * c, np.float32))
        nz = np.flatnonzero(llf)
        out[f"llf_idx_{s}"] = nz.astype(np.int32)
        out[f"llf_val_{s}"] = llf[nz]
    np.savez_compressed(HERE / "transforms.npz", **out)

    img = wl.synth_image(FRAME["w"], FRAME["h"], FRAME["seed"], FRAME["kind"])
    ref.use_variant("default")
    data = ref.encode_rgb8(img, FRAME["distance"], FRAME["effort"], FRAME["gaborish"], FRAME["epf"], 1)
    full_default = ref.decode_linear_f32(data, 1)
    ref.use_variant("strict")
    fr = ref.Frame(data, 1)
    d = fr.dump()
    i = d.info
    scal = {name: np.array(getattr(i, name)) for name, _ in type(i)._fields_}
    # keep only the dequant matrices of strategies this frame uses (the rest zeroed: compresses)
    used = np.unique(d.ac_strategy[(d.ac_strategy & 1) == 1] >> 1)
    keep = np.zeros(d.dequant.size, bool)
    for s_ in used:
        n = 64 * abi.COVERED_X[s_] * abi.COVERED_Y[s_]
        for c in range(3):
            keep[d.dequant_offsets[s_, c]: d.dequant_offsets[s_, c] + n] = True
    d.dequant[~keep] = 0
    fx = dict(jxl=np.frombuffer(data, np.uint8), decoded_default=full_default,
              ac_strategy=d.ac_strategy, raw_quant=np.where(d.ac_strategy & 1, d.raw_quant, 0).astype(np.int32),
              sharpness=d.sharpness, ytox=d.ytox, ytob=d.ytob, dc=d.dc, dequant=d.dequant,
              dequant_offsets=d.dequant_offsets, coeffs=d.coeffs,
              sigma_interior=d.sigma[2:-2, 2:-2])
    for k, v in scal.items():
        fx["info_" + k] = v
    for name, mask in TAPS.items():
        img_t, _ = fr.render(mask)
        fx["tap_" + name] = img_t
    fr.close()
    np.savez_compressed(HERE / "frame_small.npz", **fx)
    for f in ("transforms.npz", "frame_small.npz"):
        print(f, (HERE / f).stat().st_size, "bytes")
    outputs()
    dc_stage()
    return 0


if __name__ == "__main__":
    sys.exit(main())
