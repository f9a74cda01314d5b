#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: compile the product's CUDA sources for the host against the SIMT emulation shim
of this directory (cuda_runtime.h, cuda_fp16.h) into tests/emu/_build/libjxl_b200_emu.so.

Source rewriting is limited to what a C++ compiler cannot parse:
  KERNEL<<<grid, block, smem, stream>>>(args);   ->  EMU_LAUNCH((KERNEL), grid, block, smem, stream, args);
  extern __shared__ __align__(16) float fsm[];   ->  float* fsm = (float*)emu::dynamic_smem();
and the relative include of the public header.  Inline PTX is switched off inside the product source
itself by -DJXLB_HOST_EMU (see the top of jxl_kernels.cuh)."""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
CSRC = ROOT / "libjxl_b200" / "csrc"
OUT = HERE / "_build"
SO = OUT / "libjxl_b200_emu.so"
MASKS = (16, 17, 20, 21, 28, 29, 30, 31)
FUSED_MASKS = (16, 17, 20, 21, 28, 29, 30)
LAUNCH = re.compile(r"(\b\w+(?:<[^<>;]*>)?)<<<(.+?)>>>\((.*)\);")


def rewrite(text: str) -> str:
    text = LAUNCH.sub(lambda m: f"EMU_LAUNCH(({m.group(1)}), {m.group(2)}" + (f", {m.group(3)}" if m.group(3).strip() else "") + ");",
                      text)
    text = text.replace("extern __shared__ __align__(16) float fsm[];", "float* fsm = (float*)emu::dynamic_smem();")
    text = text.replace('#include "../../include/jxl_b200.h"', '#include "jxl_b200.h"')
    assert "<<<" not in text and "extern __shared__" not in text
    return text


def build(force: bool = False, sanitize: str = "") -> Path:
    """sanitize: "" | "address" | "thread" -> a separate library built with -fsanitize=<...> (a memcheck /
    racecheck substitute: CUDA threads are OS threads here, shared memory and "device" buffers are host
    memory).  Load it in a process started with the matching runtime preloaded (see tests/emu/README.md)."""
    global OUT, SO
    if sanitize:
        OUT = HERE / f"_build_{sanitize}"
        SO = OUT / "libjxl_b200_emu.so"
    srcs = sorted(CSRC.glob("*")) + [HERE / "cuda_runtime.h", HERE / "cuda_fp16.h", Path(__file__)]
    if not force and SO.exists() and all(SO.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return SO
    gen = OUT / "csrc"
    gen.mkdir(parents=True, exist_ok=True)
    for f in CSRC.glob("*"):
        (gen / (f.stem + ".cc" if f.suffix == ".cu" else f.name)).write_text(rewrite(f.read_text()))
    flags = ["-std=c++20", "-O1", "-g0", "-ffp-contract=off", "-fPIC", "-pthread", "-fvisibility=hidden", "-w",
             "-DJXLB_HOST_EMU=1", f"-I{HERE}", f"-I{ROOT / 'include'}", f"-I{gen}"]
    if sanitize:
        flags += [f"-fsanitize={sanitize}", "-g", "-fno-omit-frame-pointer"]
    if sanitize == "thread":
        flags += ["-DJXLB_EMU_CLAMP_GARBAGE_LANES=1"]   # see filter_strip_body in jxl_kernels.cuh
    units = [(gen / "jxl_b200.cc", OUT / "jxl_b200.o", [])]
    units += [(gen / "jxl_strip_inst.cc", OUT / f"strip_{m}.o", [f"-DSTRIP_MASK={m}"]) for m in MASKS]
    units += [(gen / "jxl_fused_inst.cc", OUT / f"fused_{m}.o", [f"-DFUSED_MASK={m}"]) for m in FUSED_MASKS]

    def cc(u):
        src, obj, defs = u
        r = subprocess.run(["g++", *flags, *defs, "-c", str(src), "-o", str(obj)], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"g++ failed for {src.name} {defs}:\n{r.stderr[-6000:]}")

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        list(ex.map(cc, units))
    subprocess.check_call(["g++", "-shared", "-pthread", *([f"-fsanitize={sanitize}"] if sanitize else []),
                           *[str(u[1]) for u in units], "-o", str(SO)])
    return SO


if __name__ == "__main__":
    san = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--sanitize=")), "")
    print(build(force="--force" in sys.argv, sanitize=san))
