"""-m gpu: the optional DC stage on the device -- quantised DC in, DequantDC +
AdaptiveDCSmoothing run by dc_dequant_kernel / dc_smooth_kernel -- against the oracle and against the
host-prepared-DC path.  (Sorted last on purpose: newest row of the scope table.)"""
import numpy as np
import pytest

import jxl_workload as wl
from libjxl_b200 import pipeline
from tests import support

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("built")]

# Written after round 1's GPU budget was spent: not yet run on hardware, but the same product source runs
# these very scenarios bit-exactly under the SIMT emulation of tests/emu (tests/test_emulated_cuda.py).


def with_quant_dc(desc, smoothing=1):
    from oracle import cpu
    yb, xb = desc.ysize_blocks, desc.xsize_blocks
    q = support.dc_stage_input(xb, yb)
    gm = np.array([[1.0, 0.5], [0.25, 0.125]], np.float32)[:(yb + 255) // 256, :(xb + 255) // 256].copy()
    # what the host would have prepared: DequantDC per DC group, then smoothing of the whole image
    dc = np.zeros((3, yb, xb), np.float32)
    for gy in range(gm.shape[0]):
        for gx in range(gm.shape[1]):
            sl = (slice(None), slice(gy * 256, (gy + 1) * 256), slice(gx * 256, (gx + 1) * 256))
            dc[sl] = cpu.dequant_dc(q[sl], support.DC_FACTORS, float(gm[gy, gx]), support.DC_CFL)
    if smoothing:
        dc = cpu.adaptive_dc_smoothing(dc, support.DC_FACTORS)
    return q, gm, dc


@pytest.mark.parametrize("w,h,smoothing", [(520, 264, 1), (2100, 600, 1), (2100, 600, 0), (17, 9, 1)])
def test_dc_stage_on_device(w, h, smoothing):
    from oracle import cpu
    desc, coeffs = wl.synthetic_frame(w, h, seed=w + h)
    q, gm, dc = with_quant_dc(desc, smoothing)
    pipe = pipeline.TransformPipeline(device=0)
    try:
        desc.dc = dc                                   # host-prepared DC planes
        want = pipe.decode_frame(desc, coeffs)
        assert np.array_equal(want, cpu.render_frame(desc, coeffs, rcp_mode=0))
        desc.quant_dc, desc.dc_group_mul, desc.dc_smoothing = q, gm, smoothing
        desc.dc_factors, desc.dc_cfl_factors = support.DC_FACTORS, support.DC_CFL
        desc.dc = None
        got = pipe.decode_frame(desc, coeffs)          # DC stage on the device
        oracle = cpu.render_frame(desc, coeffs, rcp_mode=0)   # ... and inside the oracle
        assert np.array_equal(oracle, want), "oracle: DC stage != host-prepared DC"
        assert np.array_equal(got, want), float(np.abs(got - want).max())
    finally:
        pipe.close()


@pytest.mark.parametrize("kind,mode", [("f32", "dense"), ("f32", "sparse"), ("srgb8", "sparse")])
def test_cpp_host_example_matches_python_host(tmp_path, kind, mode):
    """examples/host_feed.cc (C++ worker threads, shuffled group order, streamed output) produces the
    same bytes as the Python mirror for the same frame."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    from libjxl_b200 import abi
    from tests.test_abi import _build_host_feed
    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root / "examples"))
    import dump_frame
    exe = _build_host_feed(tmp_path)
    desc, coeffs = wl.synthetic_frame(1100, 777, seed=1877)
    if kind == "srgb8":
        desc.out_format, desc.stage_mask = abi.OUT_RGB_U8, abi.STAGE_SRGB
    dump, raw = tmp_path / "frame.bin", tmp_path / "out.raw"
    dump_frame.write_dump(dump, desc, coeffs)
    env = dict(os.environ, LD_LIBRARY_PATH=str(root / "libjxl_b200"))
    out = subprocess.run([str(exe), str(dump), str(raw), "6", mode], env=env, capture_output=True, text=True)
    assert out.returncode == 0, (out.stdout, out.stderr)
    got = np.fromfile(raw, desc.out_dtype).reshape(desc.out_shape())
    pipe = pipeline.TransformPipeline(device=0)
    try:
        want = pipe.decode_frame(desc, coeffs)
    finally:
        pipe.close()
    assert np.array_equal(got, want)
