"""Workloads for tests and bench.py: synthetic images, reference-made frames, synthetic frames.

Two kinds of frame inputs:
  * `reference_frame(...)`: a seeded synthetic image is encoded by the reference encoder and
    opened with the reference's own FrameDecoder (oracle/ref_harness.cc), which plays the host
    libjxl here: bitstream -> entropy-decoded coefficient groups + side info.  That hand-off is
    what the C ABI receives in a real integration; it is workload
    preparation, untimed, and not on the product path.
  * `synthetic_frame(...)`: no reference needed; random coefficients and a strategy map that
    cycles through all 27 AcStrategy types (cjxl never emits > 64x64, SURVEY.md §8c caveat).
"""
from __future__ import annotations

import hashlib
import os
import pickle
from pathlib import Path

import numpy as np

from libjxl_b200 import abi

ROOT = Path(__file__).resolve().parent
CACHE = Path(os.environ.get("JXL_B200_CACHE", "/tmp/jxl_b200_cache"))


def synth_image(w: int, h: int, seed: int = 1234, kind: str = "photo") -> np.ndarray:
    """Seeded RGB8 test image (low-frequency gradients + 4x4-block noise
    sigma 12 + pixel noise sigma 4 + two hard edges; `noise` = uniform noise)."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = (128 + 60 * np.sin(xx / 37.0)[..., None] * np.array([1, 0.7, 0.4], np.float32)
           + 40 * np.cos(yy / 23.0)[..., None] * np.array([0.3, 1, 0.6], np.float32))
    if kind == "photo":
        blk = rng.normal(0, 12, ((h + 3) // 4, (w + 3) // 4, 3)).astype(np.float32)
        img += np.repeat(np.repeat(blk, 4, 0), 4, 1)[:h, :w]
        img += rng.normal(0, 4, (h, w, 3)).astype(np.float32)
    else:  # "smooth": large transforms dominate
        img += 20 * np.sin((xx + 2 * yy) / 91.0)[..., None]
        img += rng.normal(0, 0.7, (h, w, 3)).astype(np.float32)
    img[h // 5: 2 * h // 5, 3 * w // 5: 4 * w // 5] += 80
    img[3 * h // 5: 4 * h // 5, w // 6: w // 3] -= 70
    return np.clip(img, 0, 255).astype(np.uint8)


def strategy_histogram(ac_strategy: np.ndarray) -> dict[str, int]:
    first = ac_strategy[(ac_strategy & 1) == 1] >> 1
    cnt = np.bincount(first, minlength=27)
    return {abi.STRATEGY_NAMES[i]: int(c) for i, c in enumerate(cnt) if c}


def reference_frame(w: int, h: int, distance: float = 1.0, effort: int = 7, gaborish: int = -1,
                    epf: int = -1, seed: int = 1234, kind: str = "photo", threads: int | None = None,
                    cache: bool = True, want_decoded: bool = True):
    """Returns dict(desc, coeffs, jxl, decoded, hist, bpp). Uses oracle/_ref (the reference)."""
    from oracle import cpu as ocpu
    from oracle import ref
    key = hashlib.sha1(f"v2-{w}x{h}-d{distance}-e{effort}-g{gaborish}-p{epf}-s{seed}-{kind}".encode()).hexdigest()[:16]
    path = CACHE / f"frame_{key}.pkl"
    if cache and path.exists():
        with open(path, "rb") as f:
            out = pickle.load(f)
        out["desc"] = ocpu.desc_from_dump(out.pop("dump"))
        return out
    img = synth_image(w, h, seed, kind)
    threads = threads or os.cpu_count() or 1
    data = ref.encode_rgb8(img, distance, effort, gaborish, epf, threads)
    fr = ref.Frame(data, threads)
    d = fr.dump()
    fr.close()
    out = {"coeffs": d.coeffs, "jxl": data, "decoded": d.decoded if want_decoded else None,
           "hist": strategy_histogram(d.ac_strategy), "bpp": 8.0 * len(data) / (w * h)}
    if cache:
        CACHE.mkdir(parents=True, exist_ok=True)
        dd = _DumpLite(d)
        with open(path, "wb") as f:
            pickle.dump({**out, "dump": dd}, f, protocol=4)
    out["desc"] = ocpu.desc_from_dump(d)
    return out


class _InfoLite:
    pass


class _DumpLite:
    """Picklable copy of oracle.ref.FrameDump (ctypes structs do not pickle)."""

    def __init__(self, d):
        self.info = _InfoLite()
        for name, _ in type(d.info)._fields_:
            v = getattr(d.info, name)
            setattr(self.info, name, list(v) if hasattr(v, "__len__") else v)
        for k in ("ac_strategy", "raw_quant", "sharpness", "ytox", "ytob", "dc", "sigma", "dequant",
                  "dequant_offsets"):
            setattr(self, k, getattr(d, k))


def synthetic_dequant_table(rng) -> tuple[np.ndarray, np.ndarray]:
    """A smooth, strictly positive stand-in for DequantMatrices (values only need to be
    plausible: parity tests compare the CUDA path and the oracle on the SAME table)."""
    offs = np.zeros((27, 3), np.uint32)
    parts = []
    pos = 0
    for s in range(27):
        r, c = abi.COVERED_Y[s] * 8, abi.COVERED_X[s] * 8
        n = r * c
        lo, hi = min(r, c), max(r, c)
        ky, kx = np.mgrid[0:lo, 0:hi].astype(np.float32)
        rad = np.sqrt((ky / lo) ** 2 + (kx / hi) ** 2)
        for ch in range(3):
            m = (0.00012 + 0.0012 * rad ** 1.5) * (1.0 + 0.1 * ch) * (1 + 0.05 * rng.random((lo, hi), dtype=np.float32))
            offs[s, ch] = pos
            parts.append(m.astype(np.float32).ravel())
            pos += n
    return np.concatenate(parts), offs


def synthetic_frame(w: int, h: int, seed: int = 7, strategies: str = "all", gab: int = 1,
                    epf_iters: int = 3, ac_type: int = abi.AC_INT16, density: float = 0.15):
    """Random frame covering every AcStrategy. Returns (desc, coeffs (3, num_groups, 65536))."""
    rng = np.random.default_rng(seed)
    xb, yb = (w + 7) // 8, (h + 7) // 8
    xg, yg = (xb + 31) // 32, (yb + 31) // 32
    acs = np.zeros((yb, xb), np.uint8)
    filled = np.zeros((yb, xb), bool)
    order = list(range(27)) if strategies == "all" else [int(s) for s in strategies.split(",")]
    # Every group starts with one "primary" multi-block transform at its origin (cycling through
    # all of them, largest first), the rest is filled round-robin with whatever fits, aligned to
    # the transform's own size.
    big_first = sorted(order, key=lambda s: -abi.COVERED_X[s] * abi.COVERED_Y[s])
    multi = [s for s in big_first if abi.COVERED_X[s] * abi.COVERED_Y[s] > 1] or big_first
    pick = 0

    def place(s, x, y):
        cx, cy = abi.COVERED_X[s], abi.COVERED_Y[s]
        acs[y:y + cy, x:x + cx] = s << 1
        acs[y, x] |= 1
        filled[y:y + cy, x:x + cx] = True

    for gy in range(yg):
        for gx in range(xg):
            x0, y0 = gx * 32, gy * 32
            nbx, nby = min(32, xb - x0), min(32, yb - y0)
            for k in range(len(multi)):
                s = multi[(gy * xg + gx + k) % len(multi)]
                if abi.COVERED_X[s] <= nbx and abi.COVERED_Y[s] <= nby:
                    place(s, x0, y0)
                    break
            for by in range(nby):
                for bx in range(nbx):
                    if filled[y0 + by, x0 + bx]:
                        continue
                    for attempt in range(len(order)):
                        s = order[(pick + attempt) % len(order)]
                        cx, cy = abi.COVERED_X[s], abi.COVERED_Y[s]
                        if bx % cx or by % cy or bx + cx > nbx or by + cy > nby:
                            continue
                        if filled[y0 + by:y0 + by + cy, x0 + bx:x0 + bx + cx].any():
                            continue
                        place(s, x0 + bx, y0 + by)
                        pick += 1 + attempt
                        break
                    else:
                        raise AssertionError("no strategy fits")
    first = (acs & 1) == 1
    quant = np.where(first, rng.integers(1, 257, (yb, xb)), 0).astype(np.int32)
    dc = rng.normal(0, 0.08, (3, yb, xb)).astype(np.float32)
    dc[1] += 0.35
    dc[2] += 0.3
    cm = ((yb + 7) // 8, (xb + 7) // 8)
    ytox = rng.integers(-20, 21, cm).astype(np.int8)
    ytob = rng.integers(-20, 21, cm).astype(np.int8)
    sharp = rng.integers(0, 8, (yb, xb)).astype(np.uint8)
    dq, offs = synthetic_dequant_table(rng)
    m, bias, cbrt = abi.default_opsin()
    desc = abi.FrameDesc(
        xsize=w, ysize=h, ac_strategy=acs, raw_quant=quant, dc=dc, ytox=ytox, ytob=ytob, dequant=dq,
        dequant_offsets=offs, inv_global_scale=65536.0 / 24000.0, quant_scale=24000.0 / 65536.0,
        x_dm_multiplier=0.8, b_dm_multiplier=1.25, gab=gab, epf_iters=epf_iters, epf_sharpness=sharp,
        inverse_opsin_matrix=m, opsin_biases=bias, opsin_biases_cbrt=cbrt, ac_type=ac_type)
    dt = np.int16 if ac_type == abi.AC_INT16 else np.int32
    coeffs = np.zeros((3, xg * yg, 65536), dt)
    for g in range(xg * yg):
        n = desc.group_ncoeff(g)
        lap = rng.laplace(0, 2.0, (3, n))
        keep = rng.random((3, n)) < density
        q = np.where(keep, np.rint(lap), 0)
        big = rng.random((3, n)) < 0.002
        q = np.where(big, rng.integers(-300, 301, (3, n)), q)
        coeffs[:, g, :n] = q.astype(dt)
    return desc, coeffs
