#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: run a set of decode scenarios through the emulated library built with
-fsanitize=address or -fsanitize=thread (tests/emu/build_emu.py --sanitize=...).  The process has to be
started with the sanitizer runtime preloaded:

    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \\
        python tests/emu/run_sanitized.py address
    LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS=report_signal_unsafe=0 \\
        python tests/emu/run_sanitized.py thread [scenario ...]

CUDA threads are OS threads here and shared / "device" memory is host memory, so AddressSanitizer sees
out-of-bounds accesses of kernels (a memcheck substitute) and ThreadSanitizer sees missing barriers
(a racecheck substitute)."""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import jxl_workload as wl  # noqa: E402
from libjxl_b200 import abi, pipeline  # noqa: E402
from tests.emu import build_emu  # noqa: E402


def scenarios(pipe):
    def all27():
        desc, coeffs = wl.synthetic_frame(520, 264, seed=1)
        pipe.decode_frame(desc, coeffs)

    def int32_small():
        desc, coeffs = wl.synthetic_frame(300, 200, seed=2, ac_type=abi.AC_INT32)
        pipe.decode_frame(desc, coeffs)

    def chains():
        for gab in (0, 1):
            for epf in (0, 1, 2, 3):
                desc, coeffs = wl.synthetic_frame(150, 140, seed=gab * 4 + epf, gab=gab, epf_iters=epf)
                pipe.decode_frame(desc, coeffs)

    def tile_kernel_and_formats():
        for fmt in range(6):
            desc, coeffs = wl.synthetic_frame(131, 77, seed=fmt)
            desc.out_format, desc.stage_mask = fmt, abi.STAGE_SRGB
            pipe.decode_frame(desc, coeffs)
            desc.stage_mask = abi.STAGE_EXPLICIT | 1 | 8 | 16 | abi.STAGE_SRGB
            pipe.decode_frame(desc, coeffs)

    def sparse():
        desc, coeffs = wl.synthetic_frame(300, 300, seed=5, ac_type=abi.AC_INT32)
        coeffs = coeffs.copy()
        coeffs[1, 2, 100] = 200000
        pipe.decode_frame(desc, coeffs, sparse=True, order=[3, 1, 0, 2], stream_output=True)

    def ragged():
        for w, h in ((1, 1), (9, 17), (255, 257)):
            desc, coeffs = wl.synthetic_frame(w, h, seed=w)
            pipe.decode_frame(desc, coeffs)

    def replicas():
        desc, coeffs = wl.synthetic_frame(301, 203, seed=7)
        desc.out_format, desc.stage_mask = abi.OUT_RGB_U8, abi.STAGE_SRGB
        dev = np.ascontiguousarray(coeffs)
        pipe.set_device_coefficients([dev[c].ctypes.data for c in range(3)])
        pipe.frame_begin(desc)
        n = desc.ysize * desc.out_row_bytes
        bufs = [np.zeros(n + 16, np.uint8) for _ in range(3)]
        base = lambda a: a.ctypes.data + (-a.ctypes.data) % 8
        pipe.set_output_replicas([base(b) for b in bufs[1:]])
        pipe.render_device(base(bufs[0]), desc.out_row_bytes)
        pipe.set_output_replicas([])
        pipe.set_device_coefficients(None)

    # ---- round 2 ----
    def epf_permutation():   # EPF engaged on a random subset of the blocks: the strip kernel's block permutation
        rng = np.random.default_rng(3)
        for (w, h, epf) in ((523, 90, 1), (250, 41, 3), (8, 200, 2)):
            desc, coeffs = wl.synthetic_frame(w, h, seed=w, gab=1, epf_iters=epf, density=0.05)
            desc.raw_quant = np.where(desc.raw_quant > 0, 1, 0).astype(np.int32)
            desc.epf_sharpness = np.where(rng.random(desc.epf_sharpness.shape) < 0.4, 7, 0).astype(np.uint8)
            pipe.decode_frame(desc, np.clip(coeffs, -2, 2))

    def large_transforms():  # all nine 64..256 transforms: sliced row / column passes, TMA-staged 8x8 class
        desc, coeffs = wl.synthetic_frame(1024, 520, seed=5)
        pipe.decode_frame(desc, coeffs)

    def upsampling_and_noise():
        wts = np.load(ROOT / "tests" / "golden" / "upsampling_weights.npz")
        for n, w, h in ((2, 121, 67), (8, 40, 30), (1, 301, 260)):
            desc, coeffs = wl.synthetic_frame(w, h, seed=n, epf_iters=1)
            if n > 1:
                desc.upsampling, desc.upsampling_weights = n, wts[f"weights{n}"]
                desc.xsize_upsampled, desc.ysize_upsampled = n * w - (n - 1), n * h - 1
            desc.noise, desc.noise_lut = 1, (0.001, 0.0068, 0.0039, 0.0049, 0.0059, 0.0078, 0.0088, 0.0107)
            desc.out_format, desc.stage_mask = abi.OUT_RGB_U8, abi.STAGE_SRGB
            pipe.decode_frame(desc, coeffs)

    def gather_chunks():     # copy-engine / SM-kernel gather of row chunks (JXLGPU_GATHER read per replica set)
        import os
        for mode in ("", "sm"):
            os.environ["JXLGPU_GATHER"] = mode
            desc, coeffs = wl.synthetic_frame(304, 520, seed=9, epf_iters=1)
            dev = np.ascontiguousarray(coeffs)
            pipe.set_device_coefficients([dev[c].ctypes.data for c in range(3)])
            pipe.frame_begin(desc)
            n = desc.ysize * desc.out_row_bytes
            bufs = [np.zeros(n + 32, np.uint8) for _ in range(3)]
            base = lambda a: a.ctypes.data + (-a.ctypes.data) % 16
            pipe.set_output_replicas([base(b) for b in bufs[1:]])
            pipe.render_device(base(bufs[0]), desc.out_row_bytes)
            pipe.synchronize()
            pipe.set_output_replicas([])
            pipe.set_device_coefficients(None)
        os.environ.pop("JXLGPU_GATHER", None)

    return dict(all27=all27, int32_small=int32_small, chains=chains, tile_kernel_and_formats=tile_kernel_and_formats,
                sparse=sparse, ragged=ragged, replicas=replicas, epf_permutation=epf_permutation,
                large_transforms=large_transforms, upsampling_and_noise=upsampling_and_noise, gather_chunks=gather_chunks)


def main() -> int:
    kind = sys.argv[1]
    so = build_emu.build(sanitize=kind)
    pipeline._lib = pipeline.bind(C.CDLL(str(so)))
    pipe = pipeline.TransformPipeline(device=0, num_host_threads=2)
    sc = scenarios(pipe)
    for name in (sys.argv[2:] or list(sc)):
        print(f"[{kind}] {name}", flush=True)
        sc[name]()
    pipe.close()
    print(f"[{kind}] done: no report above means clean", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
