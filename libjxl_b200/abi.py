"""ctypes mirror of include/jxl_b200.h (the C ABI) + the host-side frame description.

`FrameDesc` is the Python view of what libjxl holds in PassesSharedState /
PassesDecoderState when DecodeGroup runs (lib/jxl/passes_state.h:48-96,
lib/jxl/dec_cache.h:86-188): numpy planes + scalars. `FrameDesc.to_struct()` pins
them into a `jxlgpu_frame`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

ABI_VERSION = 6
NUM_STRATEGIES = 27
GROUP_DIM = 256
GROUP_COEFFS = 65536

OK, ERR_INVALID_ARGUMENT, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA, ERR_OOM, ERR_STATE = range(7)
AC_INT16, AC_INT32 = 0, 1
OUT_RGB_F32, OUT_PLANAR_F32, OUT_RGB_U8, OUT_RGBA_U8, OUT_RGB_U16, OUT_RGB_F16 = range(6)
STAGE_GAB, STAGE_EPF0, STAGE_EPF1, STAGE_EPF2, STAGE_XYB, STAGE_SRGB = 1, 2, 4, 8, 16, 32
# out_format -> (numpy dtype, channels per pixel); OUT_PLANAR_F32 is the one planar layout
OUT_LAYOUT = {OUT_RGB_F32: ("float32", 3), OUT_PLANAR_F32: ("float32", 1), OUT_RGB_U8: ("uint8", 3),
              OUT_RGBA_U8: ("uint8", 4), OUT_RGB_U16: ("uint16", 3), OUT_RGB_F16: ("float16", 3)}
STAGE_EXPLICIT = 1 << 31

# AcStrategy geometry, lib/jxl/ac_strategy.h:148-173
COVERED_X = (1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32)
COVERED_Y = (1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16)
STRATEGY_NAMES = ("DCT", "IDENTITY", "DCT2X2", "DCT4X4", "DCT16X16", "DCT32X32", "DCT16X8", "DCT8X16",
                  "DCT32X8", "DCT8X32", "DCT32X16", "DCT16X32", "DCT4X8", "DCT8X4", "AFV0", "AFV1",
                  "AFV2", "AFV3", "DCT64X64", "DCT64X32", "DCT32X64", "DCT128X128", "DCT128X64",
                  "DCT64X128", "DCT256X256", "DCT256X128", "DCT128X256")


class JxlGpuConfig(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32),
                ("num_host_threads", C.c_uint32), ("flags", C.c_uint32)]


class JxlGpuFrame(C.Structure):
    _fields_ = [
        ("xsize", C.c_uint32), ("ysize", C.c_uint32),
        ("xsize_blocks", C.c_uint32), ("ysize_blocks", C.c_uint32),
        ("ac_type", C.c_uint32),
        ("band_y0_groups", C.c_uint32), ("band_ny_groups", C.c_uint32),
        ("ac_strategy", C.c_void_p), ("ac_strategy_stride", C.c_size_t),
        ("raw_quant", C.c_void_p), ("raw_quant_stride", C.c_size_t),
        ("epf_sharpness", C.c_void_p), ("epf_sharpness_stride", C.c_size_t),
        ("ytox_map", C.c_void_p), ("ytob_map", C.c_void_p), ("cmap_stride", C.c_size_t),
        ("dc", C.c_void_p * 3), ("dc_stride", C.c_size_t),
        ("dequant_table", C.c_void_p), ("dequant_table_floats", C.c_size_t),
        ("dequant_offsets", C.c_uint32 * (3 * NUM_STRATEGIES)),
        ("inv_global_scale", C.c_float), ("quant_scale", C.c_float),
        ("x_dm_multiplier", C.c_float), ("b_dm_multiplier", C.c_float),
        ("quant_biases", C.c_float * 4),
        ("cfl_base_x", C.c_float), ("cfl_base_b", C.c_float), ("cfl_color_scale", C.c_float),
        ("gab", C.c_uint32), ("gab_weights", C.c_float * 6),
        ("epf_iters", C.c_uint32), ("epf_sharp_lut", C.c_float * 8),
        ("epf_channel_scale", C.c_float * 3),
        ("epf_quant_mul", C.c_float), ("epf_pass0_sigma_scale", C.c_float),
        ("epf_pass2_sigma_scale", C.c_float), ("epf_border_sad_mul", C.c_float),
        ("inverse_opsin_matrix", C.c_float * 9),
        ("opsin_biases", C.c_float * 3), ("opsin_biases_cbrt", C.c_float * 3),
        ("out_format", C.c_uint32), ("stage_mask", C.c_uint32),
        ("quant_dc", C.c_void_p * 3), ("quant_dc_stride", C.c_size_t),
        ("dc_factors", C.c_float * 3), ("dc_cfl_factors", C.c_float * 3),
        ("dc_group_mul", C.c_void_p), ("dc_smoothing", C.c_uint32),
        ("upsampling", C.c_uint32), ("xsize_upsampled", C.c_uint32), ("ysize_upsampled", C.c_uint32),
        ("upsampling_weights", C.c_void_p),
        ("noise", C.c_uint32), ("noise_lut", C.c_float * 8),
        ("visible_frame_index", C.c_uint32), ("nonvisible_frame_index", C.c_uint32),
        ("color_transform", C.c_uint32), ("reserved1", C.c_uint32),
    ]


def _f32(a, n=None):
    a = np.ascontiguousarray(a, np.float32).ravel()
    assert n is None or a.size == n, (a.size, n)
    return a


class JxlGpuSparseGroup(C.Structure):
    """jxlgpu_sparse_group (include/jxl_b200.h)."""
    _fields_ = [("group_idx", C.c_uint32), ("n16", C.c_uint32 * 3), ("n32", C.c_uint32 * 3),
                ("nz16", C.c_void_p * 3), ("nz32", C.c_void_p * 3)]


def pack_sparse(plane: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """One (group, channel) plane of quantised coefficients -> (nz16 words, nz32 word pairs), the two
    lists of jxlgpu_sparse_group -- what an entropy decoder that appends instead of scattering emits."""
    idx = np.flatnonzero(plane)
    v = plane[idx].astype(np.int64)
    small = (v >= -32768) & (v <= 32767)
    w16 = ((idx[small].astype(np.uint32) << np.uint32(16)) | (v[small] & 0xffff).astype(np.uint32)).astype(np.uint32)
    big = np.empty((int((~small).sum()), 2), np.uint32)
    big[:, 0] = idx[~small]
    big[:, 1] = (v[~small] & 0xffffffff).astype(np.uint32)
    return w16, big.ravel()


@dataclass
class FrameDesc:
    """Host-side description of one VarDCT frame's hot-path inputs (see module doc)."""
    xsize: int
    ysize: int
    ac_strategy: np.ndarray          # u8  (yb, xb): (type << 1) | is_first
    raw_quant: np.ndarray            # i32 (yb, xb)
    dc: np.ndarray                   # f32 (3, yb, xb)
    ytox: np.ndarray                 # i8  (ceil(yb/8), ceil(xb/8))
    ytob: np.ndarray                 # i8
    dequant: np.ndarray              # f32 (table,)
    dequant_offsets: np.ndarray      # u32 (27, 3)
    inv_global_scale: float
    quant_scale: float
    x_dm_multiplier: float = 1.0
    b_dm_multiplier: float = 1.0
    quant_biases: tuple = (1.0 - 0.05465007330715401, 1.0 - 0.07005449891748593,
                           1.0 - 0.049935103337343655, 0.145)     # quantizer.h:52-57
    cfl_base_x: float = 0.0
    cfl_base_b: float = 1.0
    cfl_color_scale: float = 1.0 / 84.0
    gab: int = 0
    gab_weights: tuple = (1.1 * 0.104699568, 1.1 * 0.055680538) * 3  # loop_filter.cc:30-52
    epf_iters: int = 0
    epf_sharpness: np.ndarray | None = None   # u8 (yb, xb)
    epf_sharp_lut: tuple = tuple(i / 7.0 for i in range(8))
    epf_channel_scale: tuple = (40.0, 5.0, 3.5)
    epf_quant_mul: float = 0.46
    epf_pass0_sigma_scale: float = 0.9
    epf_pass2_sigma_scale: float = 6.5
    epf_border_sad_mul: float = 2.0 / 3.0
    inverse_opsin_matrix: tuple = ()
    opsin_biases: tuple = ()
    opsin_biases_cbrt: tuple = ()
    ac_type: int = AC_INT16
    out_format: int = OUT_RGB_F32
    stage_mask: int = 0
    band_y0_groups: int = 0
    band_ny_groups: int = 0
    # optional DC stage on the device: quantised DC (3, yb, xb) int32 instead of `dc`
    quant_dc: np.ndarray | None = None
    dc_factors: tuple = (0.0, 0.0, 0.0)
    dc_cfl_factors: tuple = (0.0, 0.0, 1.0)
    dc_group_mul: np.ndarray | None = None    # f32 (ceil(yb/256), ceil(xb/256)) or None
    dc_smoothing: int = 1
    # upsampling after the filters (frame_header.upsampling): 1, 2, 4, 8; weights = the 15 / 55 / 210 floats of
    # CustomTransformData; the output has xsize_upsampled x ysize_upsampled pixels (0 = upsampling * size)
    upsampling: int = 1
    upsampling_weights: np.ndarray | None = None
    xsize_upsampled: int = 0
    ysize_upsampled: int = 0
    # noise (frame flag kNoise): NoiseParams::lut and the frame indices that seed the generator
    noise: int = 0
    noise_lut: tuple = (0.0,) * 8
    visible_frame_index: int = 1
    nonvisible_frame_index: int = 0
    color_transform: int = 0     # 0 = XYB, 1 = YCbCr (JPEG-origin 4:4:4 frames)
    _keep: list = field(default_factory=list, repr=False)

    @property
    def xsize_blocks(self) -> int:
        return (self.xsize + 7) // 8

    @property
    def ysize_blocks(self) -> int:
        return (self.ysize + 7) // 8

    @property
    def xsize_groups(self) -> int:
        return (self.xsize + GROUP_DIM - 1) // GROUP_DIM

    @property
    def ysize_groups(self) -> int:
        return (self.ysize + GROUP_DIM - 1) // GROUP_DIM

    @property
    def num_groups(self) -> int:
        return self.xsize_groups * self.ysize_groups

    @property
    def out_xsize(self) -> int:
        if self.upsampling <= 1:
            return self.xsize
        return self.xsize_upsampled or self.upsampling * self.xsize

    @property
    def out_ysize(self) -> int:
        if self.upsampling <= 1:
            return self.ysize
        return self.ysize_upsampled or self.upsampling * self.ysize

    def out_shape(self, rows: int | None = None) -> tuple:
        """Shape of the (dense) output array for `rows` pixel rows (default: this band's rows)."""
        if rows is None:
            rows = self.band_rows()[1]
        if self.upsampling > 1 and rows == self.ysize:
            rows = self.out_ysize
        if self.out_format == OUT_PLANAR_F32:
            return (3, rows, self.out_xsize)
        return (rows, self.out_xsize, OUT_LAYOUT[self.out_format][1])

    @property
    def out_dtype(self):
        return np.dtype(OUT_LAYOUT[self.out_format][0])

    @property
    def out_row_bytes(self) -> int:
        dt, ch = OUT_LAYOUT[self.out_format]
        return self.out_xsize * ch * np.dtype(dt).itemsize

    def band_rows(self) -> tuple[int, int]:
        """(first pixel row, number of pixel rows) this band renders."""
        if self.band_ny_groups == 0:
            return 0, self.ysize
        y0 = self.band_y0_groups * GROUP_DIM
        y1 = min(self.ysize, (self.band_y0_groups + self.band_ny_groups) * GROUP_DIM)
        return y0, y1 - y0

    def group_ncoeff(self, g: int) -> int:
        """Coefficients per channel the entropy decoder produces for AC group g."""
        gx, gy = g % self.xsize_groups, g // self.xsize_groups
        nbx = min(32, self.xsize_blocks - gx * 32)
        nby = min(32, self.ysize_blocks - gy * 32)
        return 64 * nbx * nby

    def to_struct(self) -> JxlGpuFrame:
        yb, xb = self.ysize_blocks, self.xsize_blocks
        s = JxlGpuFrame()
        keep = self._keep
        keep.clear()

        def pin(a, dtype, shape=None):
            a = np.ascontiguousarray(a, dtype)
            if shape is not None:
                assert a.shape == shape, (a.shape, shape)
            keep.append(a)
            return a.ctypes.data

        s.xsize, s.ysize, s.xsize_blocks, s.ysize_blocks = self.xsize, self.ysize, xb, yb
        s.ac_type = self.ac_type
        s.band_y0_groups, s.band_ny_groups = self.band_y0_groups, self.band_ny_groups
        s.ac_strategy, s.ac_strategy_stride = pin(self.ac_strategy, np.uint8, (yb, xb)), xb
        s.raw_quant, s.raw_quant_stride = pin(self.raw_quant, np.int32, (yb, xb)), xb
        if self.epf_sharpness is not None:
            s.epf_sharpness, s.epf_sharpness_stride = pin(self.epf_sharpness, np.uint8, (yb, xb)), xb
        cm = ((yb + 7) // 8, (xb + 7) // 8)
        s.ytox_map, s.ytob_map = pin(self.ytox, np.int8, cm), pin(self.ytob, np.int8, cm)
        s.cmap_stride = cm[1]
        if self.quant_dc is None:
            dc = np.ascontiguousarray(self.dc, np.float32)
            assert dc.shape == (3, yb, xb), dc.shape
            keep.append(dc)
            for c in range(3):
                s.dc[c] = dc.ctypes.data + c * yb * xb * 4
            s.dc_stride = xb
        else:
            qdc = np.ascontiguousarray(self.quant_dc, np.int32)
            assert qdc.shape == (3, yb, xb), qdc.shape
            keep.append(qdc)
            for c in range(3):
                s.quant_dc[c] = qdc.ctypes.data + c * yb * xb * 4
            s.quant_dc_stride = xb
            s.dc_factors[:] = list(_f32(self.dc_factors, 3))
            s.dc_cfl_factors[:] = list(_f32(self.dc_cfl_factors, 3))
            if self.dc_group_mul is not None:
                gm = np.ascontiguousarray(self.dc_group_mul, np.float32)
                assert gm.shape == ((yb + 255) // 256, (xb + 255) // 256), gm.shape
                keep.append(gm)
                s.dc_group_mul = gm.ctypes.data
            s.dc_smoothing = int(self.dc_smoothing)
        if self.upsampling > 1:
            nw = {2: 15, 4: 55, 8: 210}[int(self.upsampling)]
            s.upsampling = int(self.upsampling)
            s.xsize_upsampled, s.ysize_upsampled = self.out_xsize, self.out_ysize
            s.upsampling_weights = pin(np.asarray(self.upsampling_weights, np.float32).ravel()[:nw], np.float32, (nw,))
        s.color_transform = int(self.color_transform)
        if self.noise:
            s.noise = 1
            s.noise_lut[:] = list(_f32(self.noise_lut, 8))
            s.visible_frame_index, s.nonvisible_frame_index = int(self.visible_frame_index), int(self.nonvisible_frame_index)
        s.dequant_table = pin(self.dequant, np.float32)
        s.dequant_table_floats = int(np.asarray(self.dequant).size)
        offs = np.asarray(self.dequant_offsets, np.uint32).reshape(NUM_STRATEGIES * 3)
        for i, v in enumerate(offs):
            s.dequant_offsets[i] = int(v)
        s.inv_global_scale, s.quant_scale = self.inv_global_scale, self.quant_scale
        s.x_dm_multiplier, s.b_dm_multiplier = self.x_dm_multiplier, self.b_dm_multiplier
        s.quant_biases[:] = list(_f32(self.quant_biases, 4))
        s.cfl_base_x, s.cfl_base_b, s.cfl_color_scale = self.cfl_base_x, self.cfl_base_b, self.cfl_color_scale
        s.gab, s.epf_iters = int(self.gab), int(self.epf_iters)
        s.gab_weights[:] = list(_f32(self.gab_weights, 6))
        s.epf_sharp_lut[:] = list(_f32(self.epf_sharp_lut, 8))
        s.epf_channel_scale[:] = list(_f32(self.epf_channel_scale, 3))
        s.epf_quant_mul = self.epf_quant_mul
        s.epf_pass0_sigma_scale = self.epf_pass0_sigma_scale
        s.epf_pass2_sigma_scale = self.epf_pass2_sigma_scale
        s.epf_border_sad_mul = self.epf_border_sad_mul
        s.inverse_opsin_matrix[:] = list(_f32(self.inverse_opsin_matrix, 9))
        s.opsin_biases[:] = list(_f32(self.opsin_biases, 3))
        s.opsin_biases_cbrt[:] = list(_f32(self.opsin_biases_cbrt, 3))
        s.out_format, s.stage_mask = self.out_format, self.stage_mask
        return s


def default_opsin(intensity_target: float = 255.0):
    """OpsinParams::Init (lib/jxl/opsin_params.cc:35-45, lib/jxl/cms/opsin_params.h:36-62):
    spec inverse opsin absorbance matrix x 255/intensity_target, biases and their cube roots,
    computed in float32 like the reference."""
    m = np.array([[11.031566901960783, -9.866943921568629, -0.16462299647058826],
                  [-3.254147380392157, 4.418770392156863, -0.16462299647058826],
                  [-3.6588512862745097, 2.7129230470588235, 1.9459282392156863]], np.float32)
    m = (m * np.float32(255.0 / intensity_target)).astype(np.float32)
    bias = np.float32(-0.0037930732552754493)
    return tuple(m.ravel()), (bias,) * 3, (np.cbrt(bias).astype(np.float32),) * 3
