"""Host-side mirror of the C ABI (include/jxl_b200.h): the object a libjxl-like host drives.

`TransformPipeline` plays the role the render pipeline + DecodeGroup play inside
FrameDecoder (lib/jxl/dec_frame.cc:573-735): begin a frame with its side information,
submit entropy-decoded coefficient groups from worker threads, finish the frame.

There is no CPU fallback here: if libjxl_b200.so or a CUDA device is missing the
constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import abi

PKG = Path(__file__).resolve().parent
SO = PKG / "libjxl_b200.so"
CSRC = PKG / "csrc"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC,-fvisibility=hidden"]
OBJ = PKG / "build"
# the stage chains with their own row-streaming filter instantiation (one translation unit each)
STRIP_MASKS = (16, 17, 20, 21, 28, 29, 30, 31)
# the stage chains of the fused decode kernel (csrc/jxl_fused.cuh), one translation unit each
FUSED_MASKS = (16, 17, 20, 21, 28, 29, 30)

EXPORTS = ["jxlgpu_abi_version", "jxlgpu_error_string", "jxlgpu_last_error", "jxlgpu_create",
           "jxlgpu_destroy", "jxlgpu_frame_begin", "jxlgpu_frame_set_output", "jxlgpu_submit_group",
           "jxlgpu_submit_groups", "jxlgpu_submit_groups_sparse", "jxlgpu_frame_finish",
           "jxlgpu_set_device_coefficients", "jxlgpu_render_device", "jxlgpu_set_output_replicas", "jxlgpu_device_output",
           "jxlgpu_device_xyb", "jxlgpu_synchronize", "jxlgpu_launch_count", "jxlgpu_alloc_pinned",
           "jxlgpu_free_pinned", "jxlgpu_set_profiling", "jxlgpu_kernel_times"]


class JxlGpuError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: error {code}" + (f" ({detail})" if detail else ""))


def build(force: bool = False, verbose_ptxas: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into libjxl_b200.so (in-tree): the main translation unit and one
    per filter stage chain, compiled in parallel, then linked."""
    from concurrent.futures import ThreadPoolExecutor
    hdrs = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PKG.parent / "include" / "jxl_b200.h"]
    units = [(CSRC / "jxl_b200.cu", OBJ / "jxl_b200.o", [])]
    units += [(CSRC / "jxl_strip_inst.cu", OBJ / f"jxl_strip_{m}.o", [f"-DSTRIP_MASK={m}"]) for m in STRIP_MASKS]
    units += [(CSRC / "jxl_fused_inst.cu", OBJ / f"jxl_fused_{m}.o", [f"-DFUSED_MASK={m}"]) for m in FUSED_MASKS]
    newest_hdr = max(h.stat().st_mtime for h in hdrs)
    newest_src = max(newest_hdr, *(u[0].stat().st_mtime for u in units))
    if not force and SO.exists() and SO.stat().st_mtime >= newest_src:
        return SO   # up to date (the objects need not be around: only the .so travels to the GPU box)
    OBJ.mkdir(exist_ok=True)

    def compile_unit(u):
        src, obj, defs = u
        if not force and obj.exists() and obj.stat().st_mtime >= max(newest_hdr, src.stat().st_mtime):
            return False
        cmd = ["nvcc", *NVCC_FLAGS, *defs, "-c", str(src), "-o", str(obj)]
        if verbose_ptxas:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose_ptxas:
            (obj.with_suffix(".ptxas.txt")).write_text(r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed for {src.name} {defs}:\n{r.stderr[-4000:]}")
        return True

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        rebuilt = list(ex.map(compile_unit, units))
    if any(rebuilt) or not SO.exists():
        subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                               *[str(u[1]) for u in units], "-o", str(SO)])
    return SO


_lib = None


def bind(L):
    """Declare the argument types of every entry point of include/jxl_b200.h on a loaded library."""
    if True:
        L.jxlgpu_abi_version.restype = C.c_uint32
        L.jxlgpu_error_string.restype = C.c_char_p
        L.jxlgpu_error_string.argtypes = [C.c_int]
        L.jxlgpu_last_error.restype = C.c_char_p
        L.jxlgpu_last_error.argtypes = [C.c_void_p]
        L.jxlgpu_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(abi.JxlGpuConfig)]
        L.jxlgpu_destroy.argtypes = [C.c_void_p]
        L.jxlgpu_frame_begin.argtypes = [C.c_void_p, C.POINTER(abi.JxlGpuFrame)]
        L.jxlgpu_frame_set_output.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.jxlgpu_submit_group.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p * 3, C.c_size_t]
        L.jxlgpu_submit_groups.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.jxlgpu_submit_groups_sparse.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.jxlgpu_frame_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.jxlgpu_set_device_coefficients.argtypes = [C.c_void_p, C.c_void_p]
        L.jxlgpu_render_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.jxlgpu_set_output_replicas.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.jxlgpu_device_output.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.jxlgpu_device_xyb.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t)]
        L.jxlgpu_synchronize.argtypes = [C.c_void_p]
        L.jxlgpu_launch_count.restype = C.c_uint64
        L.jxlgpu_launch_count.argtypes = [C.c_void_p]
        L.jxlgpu_alloc_pinned.restype = C.c_void_p
        L.jxlgpu_alloc_pinned.argtypes = [C.c_size_t]
        L.jxlgpu_free_pinned.argtypes = [C.c_void_p]
        L.jxlgpu_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.jxlgpu_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_float * 5)]
        if L.jxlgpu_abi_version() != abi.ABI_VERSION:
            raise RuntimeError("libjxl_b200.so ABI version mismatch: rebuild")
    return L


def lib():
    global _lib
    if _lib is None:
        if not SO.exists():
            raise RuntimeError(f"{SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        _lib = bind(C.CDLL(str(SO)))
    return _lib


def pinned_array(shape, dtype) -> np.ndarray:
    """numpy array over page-locked host memory (freed with the process)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib().jxlgpu_alloc_pinned(n)
    if not p:
        raise JxlGpuError(abi.ERR_OOM, "jxlgpu_alloc_pinned")
    buf = (C.c_uint8 * n).from_address(p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def pin_side_info(desc: abi.FrameDesc) -> abi.FrameDesc:
    """Move the frame's side-information planes into page-locked memory (what a host gets by handing
    libjxl a JxlMemoryManager that allocates pinned memory): frame_begin then enqueues their uploads
    asynchronously instead of paying the driver's staging copy.  They must stay untouched until
    frame_finish."""
    for name, dt in (("ac_strategy", np.uint8), ("raw_quant", np.int32), ("epf_sharpness", np.uint8),
                     ("ytox", np.int8), ("ytob", np.int8), ("dc", np.float32), ("dequant", np.float32)):
        a = getattr(desc, name)
        if a is None:
            continue
        a = np.ascontiguousarray(a, dt)
        p = pinned_array(a.shape, dt)
        p[...] = a
        setattr(desc, name, p)
    return desc


class TransformPipeline:
    def __init__(self, device: int = 0, num_host_threads: int = 1):
        self._h = C.c_void_p()
        cfg = abi.JxlGpuConfig(abi.ABI_VERSION, device, num_host_threads, 0)
        rc = lib().jxlgpu_create(C.byref(self._h), C.byref(cfg))
        if rc:
            self._h = C.c_void_p()
            raise JxlGpuError(rc, "jxlgpu_create", lib().jxlgpu_error_string(rc).decode())
        self.desc: abi.FrameDesc | None = None
        self._struct = None

    def close(self):
        if self._h:
            lib().jxlgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, where: str):
        if rc:
            raise JxlGpuError(rc, where, lib().jxlgpu_last_error(self._h).decode() or
                              lib().jxlgpu_error_string(rc).decode())

    # ---- the three calls a host decoder makes ----
    def frame_begin(self, desc: abi.FrameDesc):
        self.desc = desc
        self._struct = desc.to_struct()
        self._check(lib().jxlgpu_frame_begin(self._h, C.byref(self._struct)), "jxlgpu_frame_begin")

    def frame_set_output(self, out: np.ndarray):
        """Announce the (ideally pinned) host output array: rows stream back as they finish."""
        d = self.desc
        assert out.shape == d.out_shape() and out.dtype == d.out_dtype and out.flags.c_contiguous
        stride = d.out_row_bytes
        self._check(lib().jxlgpu_frame_set_output(self._h, out.ctypes.data, stride), "jxlgpu_frame_set_output")

    def submit_group(self, group_idx: int, coeff_xyb, thread_id: int = 0, ncoeff: int | None = None):
        """coeff_xyb: three 1-D arrays (X, Y, B) of the frame's ac_type for AC group `group_idx`."""
        want = np.int16 if self.desc.ac_type == abi.AC_INT16 else np.int32
        arrs = [np.ascontiguousarray(a, want) if a.dtype != want or not a.flags.c_contiguous else a
                for a in coeff_xyb]
        n = ncoeff if ncoeff is not None else self.desc.group_ncoeff(group_idx)
        ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in arrs])
        self._check(lib().jxlgpu_submit_group(self._h, group_idx, thread_id, ptrs, n), "jxlgpu_submit_group")

    @staticmethod
    def make_batch(groups, host_groups):
        """Pre-marshalled arguments for submit_batch: (n, idx array, ptr array, ncoeff array)."""
        n = len(groups)
        idx = (C.c_uint32 * n)(*groups)
        ptrs = (C.c_void_p * (3 * n))(*[host_groups[g][c].ctypes.data for g in groups for c in range(3)])
        nco = (C.c_size_t * n)(*[host_groups[g][0].size for g in groups])
        return n, idx, ptrs, nco

    def submit_batch(self, batch, thread_id: int = 0):
        n, idx, ptrs, nco = batch
        self._check(lib().jxlgpu_submit_groups(self._h, n, idx, thread_id, ptrs, nco), "jxlgpu_submit_groups")

    def make_sparse_batch(self, groups, coeffs, pinned: bool = True):
        """Sparse hand-off of `groups` (jxlgpu_submit_groups_sparse): the non-zero lists of every
        (group, channel), packed back to back into one (pinned) host buffer so that the batch is one
        DMA.  coeffs: (3, num_groups, 65536).  Returns (n, struct array, buffer, bytes)."""
        lists = []
        for g in groups:
            n = self.desc.group_ncoeff(g)
            lists.append([abi.pack_sparse(coeffs[c, g, :n]) for c in range(3)])
        total = sum(a.size + b.size for gl in lists for (a, b) in gl)
        buf = pinned_array((max(total, 1),), np.uint32) if pinned else np.empty(max(total, 1), np.uint32)
        arr = (abi.JxlGpuSparseGroup * len(groups))()
        off = 0
        for i, (g, gl) in enumerate(zip(groups, lists)):
            arr[i].group_idx = int(g)
            for c, (a, b) in enumerate(gl):
                for name, cnt_name, lst, per in (("nz16", "n16", a, 1), ("nz32", "n32", b, 2)):
                    buf[off:off + lst.size] = lst
                    getattr(arr[i], cnt_name)[c] = lst.size // per
                    getattr(arr[i], name)[c] = buf.ctypes.data + 4 * off if lst.size else None
                    off += lst.size
        return len(groups), arr, buf, 4 * total

    def submit_sparse_batch(self, batch, thread_id: int = 0):
        n, arr = batch[0], batch[1]
        self._check(lib().jxlgpu_submit_groups_sparse(self._h, n, arr, thread_id), "jxlgpu_submit_groups_sparse")

    def frame_finish(self, out: np.ndarray | None = None) -> np.ndarray:
        d = self.desc
        if out is None:
            out = np.empty(d.out_shape(), d.out_dtype)
        assert out.shape == d.out_shape() and out.dtype == d.out_dtype and out.flags.c_contiguous
        stride = d.out_row_bytes
        self._check(lib().jxlgpu_frame_finish(self._h, out.ctypes.data, stride), "jxlgpu_frame_finish")
        return out

    # convenience: whole frame from a (3, num_groups, 65536) host array
    def decode_frame(self, desc: abi.FrameDesc, coeffs: np.ndarray, out: np.ndarray | None = None,
                     order=None, stream_output: bool = False, sparse: bool = False) -> np.ndarray:
        """frame_begin + submit_group for every group (`order`: submission order) + frame_finish.
        sparse: hand the groups over as non-zero lists, one AC-group row per call."""
        self.set_device_coefficients(None)
        self.frame_begin(desc)
        if stream_output:
            if out is None:
                out = np.empty(desc.out_shape(), desc.out_dtype)
            self.frame_set_output(out)
        order = list(order if order is not None else range(desc.num_groups))
        if sparse:
            keep = []
            xg = desc.xsize_groups
            for i in range(0, len(order), xg):
                keep.append(self.make_sparse_batch(order[i:i + xg], coeffs, pinned=False))
                self.submit_sparse_batch(keep[-1])
            res = self.frame_finish(out)
            del keep
            return res
        for g in order:
            self.submit_group(g, [coeffs[c, g] for c in range(3)])
        return self.frame_finish(out)

    # ---- device-resident entry points ----
    def set_device_coefficients(self, ptrs):
        if ptrs is None:
            self._check(lib().jxlgpu_set_device_coefficients(self._h, None), "jxlgpu_set_device_coefficients")
            return
        arr = (C.c_void_p * 3)(*ptrs)
        self._keep_ptrs = arr
        self._check(lib().jxlgpu_set_device_coefficients(self._h, arr), "jxlgpu_set_device_coefficients")

    def render_device(self, dev_out: int = 0, out_stride_bytes: int = 0, stream: int = 0):
        self._check(lib().jxlgpu_render_device(self._h, dev_out or None, out_stride_bytes, stream or None),
                    "jxlgpu_render_device")

    def set_output_replicas(self, ptrs, multicast_ptr: int = 0):
        """Fused all-gather: device addresses (peer-mapped) of this band's slot in every rank's frame."""
        ptrs = list(ptrs or [])
        arr = (C.c_void_p * max(1, len(ptrs)))(*ptrs) if ptrs else None
        self._check(lib().jxlgpu_set_output_replicas(self._h, len(ptrs), arr, multicast_ptr or None),
                    "jxlgpu_set_output_replicas")

    def device_output(self) -> tuple[int, int]:
        p, s = C.c_void_p(), C.c_size_t()
        self._check(lib().jxlgpu_device_output(self._h, C.byref(p), C.byref(s)), "jxlgpu_device_output")
        return p.value, s.value

    def device_xyb(self) -> tuple[int, int, int]:
        p, ps, rs = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._check(lib().jxlgpu_device_xyb(self._h, C.byref(p), C.byref(ps), C.byref(rs)), "jxlgpu_device_xyb")
        return p.value, ps.value, rs.value

    def synchronize(self):
        self._check(lib().jxlgpu_synchronize(self._h), "jxlgpu_synchronize")

    def set_profiling(self, enable: bool):
        self._check(lib().jxlgpu_set_profiling(self._h, int(enable)), "jxlgpu_set_profiling")

    def kernel_times_ms(self) -> dict[str, float]:
        ms = (C.c_float * 5)()
        self._check(lib().jxlgpu_kernel_times(self._h, C.byref(ms)), "jxlgpu_kernel_times")
        return dict(zip(("plan", "idct8", "idct_mid", "idct_large", "filter"), [float(v) for v in ms]))

    def launch_count(self) -> int:
        return int(lib().jxlgpu_launch_count(self._h))
