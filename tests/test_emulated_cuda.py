"""CPU execution of the PRODUCT's CUDA source: tests/emu compiles libjxl_b200/csrc/*.cu(h) for the host
against a SIMT emulation shim (one OS thread per CUDA thread, real barriers) and this file runs the
kernels + the C-ABI host code (scheduler, sparse hand-off, DC stage, output packing) against the oracle --
bit for bit, without a GPU.  It does not replace the -m gpu tests (no real memory system, no PTX); it
lets every round start from kernels whose logic is already known to be right.  Small frames only."""
import ctypes as C

import numpy as np
import pytest

import jxl_workload as wl
from libjxl_b200 import abi, pipeline
from tests import support


@pytest.fixture(scope="module")
def emu_pipe():
    from tests.emu import build_emu
    so = build_emu.build()
    saved = pipeline._lib
    pipeline._lib = pipeline.bind(C.CDLL(str(so)))      # the emulated library instead of libjxl_b200.so
    try:
        p = pipeline.TransformPipeline(device=0, num_host_threads=2)
        yield p
        p.close()
    finally:
        pipeline._lib = saved


def oracle(desc, coeffs):
    from oracle import cpu
    return cpu.render_frame(desc, coeffs, rcp_mode=0)


def same(a, b):
    if a.dtype == np.float16:
        a, b = a.view(np.uint16), b.view(np.uint16)
    return a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.timeout(600)
def test_emulated_golden_frame(emu_pipe):
    """tests/golden/frame_small.npz (real bitstream, Gaborish + EPF 0/1/2): strip kernel chain."""
    desc, coeffs, _ = support.golden_desc()
    assert same(emu_pipe.decode_frame(desc, coeffs), oracle(desc, coeffs))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("w,h,ac_type", [(520, 264, abi.AC_INT16), (300, 200, abi.AC_INT32)])
def test_emulated_all_strategy_frame(emu_pipe, w, h, ac_type):
    desc, coeffs = wl.synthetic_frame(w, h, seed=w + h, ac_type=ac_type)
    assert same(emu_pipe.decode_frame(desc, coeffs), oracle(desc, coeffs))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("w,h,smoothing", [(520, 264, 1), (2100, 40, 1), (2100, 40, 0), (17, 9, 1)])
def test_emulated_dc_stage(emu_pipe, w, h, smoothing):
    """dc_dequant_kernel + dc_smooth_kernel + the frame_begin plumbing (quantised DC in, two DC groups
    per row at 2100 px) == host-prepared DC planes == the oracle's own DC stage."""
    from tests.test_zz_dc_stage_gpu import with_quant_dc
    desc, coeffs = wl.synthetic_frame(w, h, seed=w + h)
    q, gm, dc = with_quant_dc(desc, smoothing)
    desc.dc = dc
    want = emu_pipe.decode_frame(desc, coeffs)
    assert same(want, oracle(desc, coeffs))
    desc.quant_dc, desc.dc_group_mul, desc.dc_smoothing = q, gm, smoothing
    desc.dc_factors, desc.dc_cfl_factors = support.DC_FACTORS, support.DC_CFL
    desc.dc = None
    assert same(emu_pipe.decode_frame(desc, coeffs), want)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("ac_type", [abi.AC_INT16, abi.AC_INT32])
def test_emulated_sparse_hand_off(emu_pipe, ac_type):
    desc, coeffs = wl.synthetic_frame(300, 300, seed=77 + ac_type, ac_type=ac_type)   # 2 x 2 groups
    if ac_type == abi.AC_INT32:
        coeffs = coeffs.copy()
        rng = np.random.default_rng(2)
        for c in range(3):
            g, k = rng.integers(0, desc.num_groups, 100), rng.integers(64, 4096, 100)
            coeffs[c, g, k] = rng.integers(-300000, 300000, 100)
    want = emu_pipe.decode_frame(desc, coeffs)
    assert same(want, oracle(desc, coeffs))
    order = np.random.default_rng(9).permutation(desc.num_groups).tolist()
    assert same(emu_pipe.decode_frame(desc, coeffs, sparse=True, order=order, stream_output=True), want)
    if ac_type == abi.AC_INT16:   # planes really are re-zeroed: an all-zero frame after one with content
        assert same(emu_pipe.decode_frame(desc, np.zeros_like(coeffs), sparse=True),
                    emu_pipe.decode_frame(desc, np.zeros_like(coeffs)))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("srgb", [0, abi.STAGE_SRGB])
@pytest.mark.parametrize("fmt", [abi.OUT_RGB_F32, abi.OUT_PLANAR_F32, abi.OUT_RGB_U8, abi.OUT_RGBA_U8,
                                 abi.OUT_RGB_U16, abi.OUT_RGB_F16])
def test_emulated_output_stages(emu_pipe, fmt, srgb):
    desc, coeffs = wl.synthetic_frame(201, 131, seed=fmt * 2 + (1 if srgb else 0))
    desc.out_format, desc.stage_mask = fmt, srgb
    assert same(emu_pipe.decode_frame(desc, coeffs), oracle(desc, coeffs))          # strip kernel
    desc.stage_mask = abi.STAGE_EXPLICIT | abi.STAGE_GAB | abi.STAGE_EPF2 | abi.STAGE_XYB | srgb
    assert same(emu_pipe.decode_frame(desc, coeffs), oracle(desc, coeffs))          # tile kernel


@pytest.mark.timeout(900)
@pytest.mark.parametrize("gab", [0, 1])
@pytest.mark.parametrize("epf_iters", [0, 1, 2, 3])
def test_emulated_production_chains(emu_pipe, gab, epf_iters):
    desc, coeffs = wl.synthetic_frame(277, 300, seed=gab * 10 + epf_iters, gab=gab, epf_iters=epf_iters)
    assert same(emu_pipe.decode_frame(desc, coeffs), oracle(desc, coeffs))


@pytest.mark.timeout(900)
def test_emulated_bands_and_shuffled_streaming(emu_pipe):
    from libjxl_b200 import sharding
    desc, coeffs = wl.synthetic_frame(260, 700, seed=21)      # 2 x 3 groups
    want = emu_pipe.decode_frame(desc, coeffs)
    assert same(want, oracle(desc, coeffs))
    order = np.random.default_rng(1).permutation(desc.num_groups).tolist()
    assert same(emu_pipe.decode_frame(desc, coeffs, order=order, stream_output=True), want)
    rows = []
    for (y0, ny) in sharding.band_partition(desc.ysize_groups, 2):
        desc.band_y0_groups, desc.band_ny_groups = y0, ny
        emu_pipe.set_device_coefficients(None)
        emu_pipe.frame_begin(desc)
        for gidx in sharding.groups_needed(desc, y0, ny):
            emu_pipe.submit_group(gidx, [coeffs[c, gidx] for c in range(3)])
        rows.append(emu_pipe.frame_finish())
    desc.band_y0_groups = desc.band_ny_groups = 0
    assert same(np.concatenate(rows, axis=0), want)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("w,h", [(1, 1), (9, 17), (255, 257), (264, 72)])
def test_emulated_ragged_sizes(emu_pipe, w, h):
    desc, coeffs = wl.synthetic_frame(w, h, seed=w * 1000 + h)
    assert same(emu_pipe.decode_frame(desc, coeffs), oracle(desc, coeffs))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("kind,mode", [("f32", "dense"), ("srgb8", "sparse")])
def test_emulated_cpp_host_example(emu_pipe, tmp_path, kind, mode):
    """examples/host_feed.cc (C++ worker threads, shuffled order, streamed output) linked against the
    emulated library: same bytes as the Python mirror."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    from tests.emu import build_emu
    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root / "examples"))
    import dump_frame
    (tmp_path / "libjxl_b200.so").symlink_to(build_emu.SO)      # what -ljxl_b200 resolves to in this test
    exe = tmp_path / "host_feed"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", str(root / "include"), str(root / "examples" / "host_feed.cc"),
                           "-L", str(tmp_path), "-ljxl_b200", "-pthread", "-o", str(exe)])
    desc, coeffs = wl.synthetic_frame(300, 300, seed=1877)
    if kind == "srgb8":
        desc.out_format, desc.stage_mask = abi.OUT_RGB_U8, abi.STAGE_SRGB
    dump, raw = tmp_path / "frame.bin", tmp_path / "out.raw"
    dump_frame.write_dump(dump, desc, coeffs)
    env = dict(os.environ, LD_LIBRARY_PATH=str(tmp_path))
    out = subprocess.run([str(exe), str(dump), str(raw), "3", mode], env=env, capture_output=True, text=True)
    assert out.returncode == 0, (out.stdout, out.stderr)
    got = np.fromfile(raw, desc.out_dtype).reshape(desc.out_shape())
    assert same(got, emu_pipe.decode_frame(desc, coeffs))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("fmt,srgb", [(abi.OUT_RGB_F32, 0), (abi.OUT_RGB_U8, abi.STAGE_SRGB), (abi.OUT_RGB_U16, 0),
                                      (abi.OUT_PLANAR_F32, 0)])
def test_emulated_fused_all_gather_replay(emu_pipe, fmt, srgb):
    """REPL=1 instantiation of the strip kernel: every finished row is replayed into the `replica`
    buffers (peer GPUs' frame slots in production) with 8-byte stores plus byte head/tail pieces.  Odd
    width and packed formats make strips start at arbitrary byte offsets; the replicas must end up
    identical to the local output, and nothing outside the band may be touched."""
    desc, coeffs = wl.synthetic_frame(301, 203, seed=90 + fmt)
    desc.out_format, desc.stage_mask = fmt, srgb
    want = oracle(desc, coeffs)
    dev = np.ascontiguousarray(coeffs)                      # "device" memory is host memory here
    emu_pipe.set_device_coefficients([dev[c].ctypes.data for c in range(3)])
    emu_pipe.frame_begin(desc)
    nbytes = want.nbytes
    guard = 64
    local = np.full(nbytes + 2 * guard, 0xAB, np.uint8)
    reps = [np.full(nbytes + 2 * guard, 0xCD, np.uint8) for _ in range(3)]
    base = lambda a: a.ctypes.data + (-a.ctypes.data) % 8 + 8       # 8-byte aligned, inside the guard
    off = base(local) - local.ctypes.data
    emu_pipe.set_output_replicas([base(r) for r in reps])
    try:
        emu_pipe.render_device(base(local), desc.out_row_bytes)
    finally:
        emu_pipe.set_output_replicas([])
        emu_pipe.set_device_coefficients(None)
    flat = want.view(np.uint8).ravel()
    assert np.array_equal(local[off:off + nbytes], flat)
    for r in reps:
        o = base(r) - r.ctypes.data
        assert np.array_equal(r[o:o + nbytes], flat)
        assert (r[:o] == 0xCD).all() and (r[o + nbytes:] == 0xCD).all()


@pytest.mark.timeout(900)
def test_emulated_multicast_replay_and_its_limits(emu_pipe):
    """The multimem.st variant of the replay (emulated as a plain store to the one multicast address):
    4-byte granules for the f32 layouts; packed layouts are refused."""
    desc, coeffs = wl.synthetic_frame(203, 131, seed=123)
    want = oracle(desc, coeffs)
    dev = np.ascontiguousarray(coeffs)
    emu_pipe.set_device_coefficients([dev[c].ctypes.data for c in range(3)])
    emu_pipe.frame_begin(desc)
    local, mc = np.zeros(want.nbytes + 16, np.uint8), np.zeros(want.nbytes + 16, np.uint8)
    base = lambda a: a.ctypes.data + (-a.ctypes.data) % 8
    emu_pipe.set_output_replicas([], base(mc))
    try:
        emu_pipe.render_device(base(local), desc.out_row_bytes)
        o = base(mc) - mc.ctypes.data
        assert np.array_equal(mc[o:o + want.nbytes], want.view(np.uint8).ravel())
        desc.out_format = abi.OUT_RGB_U8
        emu_pipe.frame_begin(desc)
        with pytest.raises(pipeline.JxlGpuError):
            emu_pipe.render_device(base(local), desc.out_row_bytes)
    finally:
        emu_pipe.set_output_replicas([])
        emu_pipe.set_device_coefficients(None)
