"""Host-side logic: frame description, geometry, synthetic workloads, band sharding."""
import numpy as np
import pytest

import jxl_workload as wl
from libjxl_b200 import abi, sharding

pytestmark = pytest.mark.usefixtures("built")


def test_geometry_matches_frame_dimensions():
    # SURVEY.md §8 size table (FrameDimensions::Set, frame_dimensions.h:34-60)
    for (w, h), (xb, yb, groups) in {(512, 512): (64, 64, 4), (4096, 4096): (512, 512, 256),
                                     (7680, 4320): (960, 540, 510), (1920, 1080): (240, 135, 40)}.items():
        d = abi.FrameDesc(xsize=w, ysize=h, ac_strategy=None, raw_quant=None, dc=None, ytox=None, ytob=None,
                          dequant=None, dequant_offsets=None, inv_global_scale=1, quant_scale=1)
        assert (d.xsize_blocks, d.ysize_blocks, d.num_groups) == (xb, yb, groups)
    d = abi.FrameDesc(xsize=7680, ysize=4320, ac_strategy=None, raw_quant=None, dc=None, ytox=None, ytob=None,
                      dequant=None, dequant_offsets=None, inv_global_scale=1, quant_scale=1)
    assert d.group_ncoeff(0) == 65536
    assert d.group_ncoeff(d.num_groups - 1) == 64 * 32 * 28      # last row: 224 px


def test_synthetic_frame_covers_all_strategies_and_is_consistent():
    desc, coeffs = wl.synthetic_frame(2048, 1032, seed=3)
    hist = wl.strategy_histogram(desc.ac_strategy)
    assert len(hist) == 27
    acs = desc.ac_strategy
    covered = np.zeros(acs.shape, int)
    for by, bx in zip(*np.nonzero(acs & 1)):
        s = acs[by, bx] >> 1
        cy, cx = abi.COVERED_Y[s], abi.COVERED_X[s]
        assert by // 32 == (by + cy - 1) // 32 and bx // 32 == (bx + cx - 1) // 32  # never crosses a group
        assert (acs[by:by + cy, bx:bx + cx] >> 1 == s).all()
        covered[by:by + cy, bx:bx + cx] += 1
    assert (covered == 1).all()
    # coefficients beyond the used part of each group are zero
    for g in range(desc.num_groups):
        assert not coeffs[:, g, desc.group_ncoeff(g):].any()


def test_struct_roundtrip_keeps_scalars():
    desc, _ = wl.synthetic_frame(264, 136, seed=1)
    s = desc.to_struct()
    assert (s.xsize, s.ysize, s.xsize_blocks, s.ysize_blocks) == (264, 136, 33, 17)
    assert s.epf_iters == 3 and s.gab == 1
    assert abs(s.epf_channel_scale[0] - 40.0) < 1e-6
    assert s.dequant_table_floats == desc.dequant.size


def test_default_opsin_matches_reference_values():
    m, bias, cbrt = abi.default_opsin()
    # values printed by the reference for an sRGB 255-nit image (oracle/ref_harness.cc dump)
    want = [11.031566619873047, -9.866944313049316, -0.16462299227714539, -3.2541472911834717,
            4.4187703132629395, -0.16462299227714539, -3.658851385116577, 2.712923049926758,
            1.9459282159805298]
    assert np.allclose(m, want, rtol=0, atol=0)
    assert abs(bias[0] - -0.0037930733524262905) == 0
    assert abs(cbrt[0] - -0.15595419704914093) < 1e-8


@pytest.mark.parametrize("yg,world", [(17, 8), (16, 8), (2, 8), (5, 2), (1, 1), (64, 8)])
def test_band_partition(yg, world):
    bands = sharding.band_partition(yg, world)
    assert len(bands) == world
    assert sum(n for _, n in bands) == yg
    y = 0
    for y0, n in bands:
        assert y0 == y
        y += n
    sizes = [n for _, n in bands]
    assert max(sizes) - min(sizes) <= 1


def test_groups_needed_includes_halo_rows():
    desc, _ = wl.synthetic_frame(600, 1100, seed=2)       # 3 x 5 groups
    assert sharding.filter_halo(desc) == 7
    xg = desc.xsize_groups
    assert sharding.groups_needed(desc, 0, 5) == list(range(15))
    assert sharding.groups_needed(desc, 2, 1) == list(range(1 * xg, 4 * xg))
    assert sharding.groups_needed(desc, 0, 1) == list(range(0, 2 * xg))
    desc.gab, desc.epf_iters = 0, 0
    assert sharding.filter_halo(desc) == 0
    assert sharding.groups_needed(desc, 2, 1) == list(range(2 * xg, 3 * xg))


def test_pack_sparse_round_trip():
    """abi.pack_sparse: the two non-zero lists of jxlgpu_sparse_group re-expand to the dense plane
    (16-bit words for |v| < 2^15, {pos, value} pairs beyond)."""
    rng = np.random.default_rng(11)
    plane = np.zeros(65536, np.int32)
    idx = rng.choice(65536, 9000, replace=False)
    plane[idx] = rng.integers(-40, 41, idx.size)
    plane[idx[:50]] = rng.integers(-2**31, 2**31 - 1, 50)
    plane[idx[50:54]] = (32767, -32768, 32768, -32769)
    w16, w32 = abi.pack_sparse(plane)
    assert w16.dtype == np.uint32 and w32.dtype == np.uint32 and w32.size % 2 == 0
    back = np.zeros_like(plane)
    back[w16 >> 16] = (w16 & 0xffff).astype(np.uint16).view(np.int16)
    back[w32[0::2]] = w32[1::2].view(np.int32)
    assert np.array_equal(back, plane)
    assert w16.size + w32.size // 2 == np.count_nonzero(plane)
    assert set(np.abs(plane[w32[0::2]]).tolist()) and np.all((plane[w32[0::2]] > 32767) | (plane[w32[0::2]] < -32768))


def _build_dc_stage_host(tmp_path):
    """The product's per-block DC-stage functions (libjxl_b200/csrc/jxl_dc_stage.h -- the code the CUDA
    kernels call) compiled for the host, contraction off like the device build."""
    import ctypes as C
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    src = tmp_path / "dc_host.cc"
    src.write_text('''
#include "jxl_dc_stage.h"
extern "C" void run_dc_stage(const int32_t* q, uint32_t xb, uint32_t yb, const float* fac, float cfl_x, float cfl_b,
                             const float* gm, int smoothing, float* deq, float* out) {
  jxlb::DcStage S{};
  S.xb = xb; S.yb = yb; S.xdg = (xb + 255) / 256; S.group_mul = gm; S.cfl_x = cfl_x; S.cfl_b = cfl_b;
  const size_t n = (size_t)xb * yb;
  for (int c = 0; c < 3; c++) { S.q[c] = q + c * n; S.deq[c] = deq + c * n; S.out[c] = out + c * n; S.dc_factors[c] = fac[c]; }
  for (uint32_t y = 0; y < yb; y++) for (uint32_t x = 0; x < xb; x++) jxlb::dc_dequant_px(S, x, y);
  for (uint32_t y = 0; y < yb; y++) for (uint32_t x = 0; x < xb; x++) jxlb::dc_smooth_px(S, x, y, smoothing != 0);
}
''')
    so = tmp_path / "dc_host.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(root / "libjxl_b200" / "csrc"),
                           str(src), "-o", str(so)])
    lib = C.CDLL(str(so))
    lib.run_dc_stage.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float * 3, C.c_float, C.c_float, C.c_void_p,
                                 C.c_int, C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("xs,ys,smoothing", [(37, 21, 1), (300, 270, 1), (300, 270, 0), (2, 9, 1), (3, 3, 1)])
def test_dc_stage_device_functions_on_host(tmp_path, xs, ys, smoothing):
    """dc_dequant_px / dc_smooth_px (what dc_dequant_kernel / dc_smooth_kernel execute per block) against
    the CPU restatement of DequantDC + AdaptiveDCSmoothing, bit for bit, including the per-DC-group
    precision factor (300 blocks = two DC groups per row)."""
    import ctypes as C
    from oracle import cpu
    from tests import support
    lib = _build_dc_stage_host(tmp_path)
    q = support.dc_stage_input(xs, ys)
    gm = np.array([[1.0, 0.5], [0.25, 0.125]], np.float32)[:(ys + 255) // 256, :(xs + 255) // 256].copy()
    deq, out = np.zeros((3, ys, xs), np.float32), np.zeros((3, ys, xs), np.float32)
    lib.run_dc_stage(q.ctypes.data, xs, ys, (C.c_float * 3)(*support.DC_FACTORS), support.DC_CFL[0], support.DC_CFL[2],
                     gm.ctypes.data, smoothing, deq.ctypes.data, out.ctypes.data)
    want = np.zeros_like(deq)
    for gy in range(gm.shape[0]):
        for gx in range(gm.shape[1]):
            sl = (slice(None), slice(gy * 256, (gy + 1) * 256), slice(gx * 256, (gx + 1) * 256))
            want[sl] = cpu.dequant_dc(q[sl], support.DC_FACTORS, float(gm[gy, gx]), support.DC_CFL)
    assert np.array_equal(deq, want)
    want_sm = cpu.adaptive_dc_smoothing(want, support.DC_FACTORS) if smoothing else want
    assert np.array_equal(out, want_sm)
    if smoothing and xs > 16:
        assert (out != deq).mean() > 0.1
