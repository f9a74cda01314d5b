"""The C-ABI library: loads without a GPU, exports every symbol include/jxl_b200.h declares,
and refuses to run without a CUDA device (no silent CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from libjxl_b200 import abi, pipeline

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.usefixtures("built")


def declared_symbols():
    txt = (ROOT / "include" / "jxl_b200.h").read_text()
    return re.findall(r"JXLGPU_API\s+[\w\s\*]+?\b(jxlgpu_\w+)\s*\(", txt)


def test_every_declared_symbol_is_exported():
    names = declared_symbols()
    assert len(names) >= 14
    lib = C.CDLL(str(pipeline.SO))
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(names) == sorted(pipeline.EXPORTS)


def test_abi_version_and_error_strings():
    lib = pipeline.lib()
    assert lib.jxlgpu_abi_version() == abi.ABI_VERSION
    assert lib.jxlgpu_error_string(0) == b"ok"
    assert b"device" in lib.jxlgpu_error_string(abi.ERR_NO_DEVICE)


def test_struct_layout_matches_header():
    """sizeof(jxlgpu_frame) computed by the C compiler == ctypes mirror."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "jxl_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(jxlgpu_frame), sizeof(jxlgpu_config), sizeof(jxlgpu_sparse_group));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        (Path(td) / "t.c").write_text(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(Path(td) / "t.c"), "-o", str(Path(td) / "t")])
        out = subprocess.check_output([str(Path(td) / "t")]).split()
    assert int(out[0]) == C.sizeof(abi.JxlGpuFrame)
    assert int(out[1]) == C.sizeof(abi.JxlGpuConfig)
    assert int(out[2]) == C.sizeof(abi.JxlGpuSparseGroup)


def test_field_offsets_match_header(tmp_path):
    """offsetof() of every jxlgpu_frame field, by the C compiler, == the ctypes mirror."""
    import subprocess
    root = Path(__file__).resolve().parents[1]
    names = [n for n, _ in abi.JxlGpuFrame._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "jxl_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", offsetof(jxlgpu_frame, {n}));' for n in names) + "return 0;}\n"
    src = tmp_path / "off.c"
    src.write_text(prog)
    exe = tmp_path / "off"
    subprocess.check_call(["gcc", "-I", str(root / "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n in names:
        assert int(got[n]) == getattr(abi.JxlGpuFrame, n).offset, n


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pipeline.JxlGpuError) as e:
        pipeline.TransformPipeline(device=0)
    assert e.value.code in (abi.ERR_NO_DEVICE, abi.ERR_CUDA)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (tier rule ③)."""
    for f in (ROOT / "libjxl_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h"):
            txt = f.read_text()
            assert "oracle" not in txt.replace("oracle/jxl_oracle.c (rcp_mode 0)", ""), f


def _build_host_feed(tmp_path):
    import subprocess
    root = Path(__file__).resolve().parents[1]
    exe = tmp_path / "host_feed"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", str(root / "include"),
                           str(root / "examples" / "host_feed.cc"), "-L", str(root / "libjxl_b200"), "-ljxl_b200",
                           "-pthread", "-o", str(exe)])
    return exe


def test_cpp_host_example_builds_and_parses(tmp_path, built):
    """examples/host_feed.cc -- the C ABI driven from plain C++ worker threads, as libjxl would -- compiles
    against include/jxl_b200.h alone, links against the library, reads the frame dump correctly, and
    without a device fails loudly in jxlgpu_create (exit code 3) instead of computing anything."""
    import os
    import subprocess
    import sys
    import numpy as np
    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root / "examples"))
    import dump_frame
    import jxl_workload as wl
    exe = _build_host_feed(tmp_path)
    desc, coeffs = wl.synthetic_frame(300, 200, seed=500)
    dump = tmp_path / "frame.bin"
    dump_frame.write_dump(dump, desc, coeffs)
    env = dict(os.environ, LD_LIBRARY_PATH=str(root / "libjxl_b200"), HOST_FEED_PARSE_ONLY="1")
    out = subprocess.run([str(exe), str(dump), str(tmp_path / "out.raw")], env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    nz = sum(int(np.count_nonzero(coeffs[c, g, :desc.group_ncoeff(g)])) for g in range(desc.num_groups) for c in range(3))
    assert f"groups={desc.num_groups} " in out.stdout and f"nonzero={nz} " in out.stdout, out.stdout
    import torch
    if not torch.cuda.is_available():
        env.pop("HOST_FEED_PARSE_ONLY")
        out = subprocess.run([str(exe), str(dump), str(tmp_path / "out.raw")], env=env, capture_output=True, text=True)
        assert out.returncode == 3 and "no CUDA device" in out.stderr, (out.returncode, out.stderr)
