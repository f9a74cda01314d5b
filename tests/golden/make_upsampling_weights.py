#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: writes tests/golden/upsampling_weights.npz -- the default upsampling weights of a
JPEG XL codestream (CustomTransformData, lib/jxl/image_metadata.cc:98-214) as the reference's own decoder state
reports them for frames encoded with resampling 2 / 4 / 8, plus one small codestream per factor for the parity tests.
Needs oracle/_ref (built from /root/reference by oracle/build_ref.py); the .npz travels to the GPU box."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import jxl_workload as wl  # noqa: E402
from oracle import ref  # noqa: E402

out = {}
for rs, w, h in ((2, 560, 80), (4, 1100, 90), (8, 2100, 100)):
    img = wl.synth_image(w, h, seed=rs)
    data = ref.encode_rgb8(img, 1.0, 7, -1, -1, 4, resampling=rs)
    fr = ref.Frame(data, 2)
    i = fr.info
    assert i.upsampling == rs
    n = {2: 15, 4: 55, 8: 210}[rs]
    out[f"weights{rs}"] = np.array(list(i.upsampling_weights)[:n], np.float32)
    out[f"jxl{rs}"] = np.frombuffer(data, np.uint8)
    fr.close()
np.savez_compressed(ROOT / "tests" / "golden" / "upsampling_weights.npz", **out)
print({k: v.shape for k, v in out.items()})
