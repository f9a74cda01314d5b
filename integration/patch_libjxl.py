#!/usr/bin/env python3
"""Build-time patch of libjxl for the GPU backend: writes PATCHED COPIES of two reference files.

    python integration/patch_libjxl.py <reference root> <output dir>

Inserts the one-line hooks of integration/libjxl_gpu_backend.h (our code) at the call sites INTEGRATION.md §2
names, into copies of lib/jxl/dec_frame.cc and lib/jxl/dec_group.cc.  The copies are build products (they go to
the git-ignored oracle/_ref/patched/, next to the objects): no reference source is committed to this
repository.  Every anchor must match exactly once -- a reference that moved on fails loudly here instead of
building something subtly different.
"""
from __future__ import annotations

import sys
from pathlib import Path

HOOK_INCLUDE = '#include "libjxl_gpu_backend.h"  // jxl_b200 integration (integration/patch_libjxl.py)\n'

DEC_FRAME = [
    # the implementation of the hooks lives in this translation unit
    ('#include "lib/jxl/dec_frame.h"\n',
     '#include "lib/jxl/dec_frame.h"\n\n#define JXLB_GPU_BACKEND_IMPLEMENTATION 1\n' + HOOK_INCLUDE +
     'extern "C" unsigned long long jxlb_gpu_backend_frames_taken(void) { return jxlb_integration::FramesTaken(); }\n'),
    # site 1 (dec_frame.cc:417-434): coefficient storage -> pinned, group-major, accumulate mode
    ('    if (store) {\n      dec_state_->coefficients->ZeroFill();\n    }\n',
     '    if (store) {\n      dec_state_->coefficients->ZeroFill();\n    }\n'
     '    jxlb_integration::WantFrame(frame_header_, dec_state_, decoded_, use_16_bit, frame_dim_.num_groups);\n'),
    # site 2 (dec_frame.cc:693-696): side information is final -> frame_begin
    ('    JXL_RETURN_IF_ERROR(ProcessACGlobal(sections[ac_global_sec].br));\n'
     '    section_status[ac_global_sec] = SectionStatus::kDone;\n',
     '    JXL_RETURN_IF_ERROR(ProcessACGlobal(sections[ac_global_sec].br));\n'
     '    section_status[ac_global_sec] = SectionStatus::kDone;\n'
     '    JXL_RETURN_IF_ERROR(jxlb_integration::BeginFrame(frame_header_, dec_state_));\n'),
    # site 3a (dec_frame.cc:503-505): zero-fill this group's pinned blocks / arm the thread's sparse sink
    ('    JXL_RETURN_IF_ERROR(DecodeGroup(\n        frame_header_, br.data(), num_passes, ac_group_id, dec_state_,\n',
     '    jxlb_integration::PrepareGroup(dec_state_, ac_group_id, decoded_passes_per_ac_group_[ac_group_id]);\n'
     '    JXL_RETURN_IF_ERROR(DecodeGroup(\n        frame_header_, br.data(), num_passes, ac_group_id, dec_state_,\n'),
    # site 3 (dec_frame.cc:506-516): the group was entropy-decoded only (DecodeGroup saw kDontDraw) -> submit
    ('        force_draw, dc_only, &should_run_pipeline));\n  }\n',
     '        force_draw, dc_only, &should_run_pipeline));\n'
     '    if (num_passes > 0 && jxlb_integration::DontDraw(dec_state_) &&\n'
     '        decoded_passes_per_ac_group_[ac_group_id] + num_passes >= frame_header_.passes.num_passes) {\n'
     '      JXL_RETURN_IF_ERROR(jxlb_integration::GroupDecoded(dec_state_, ac_group_id));\n'
     '    }\n  }\n'),
    # site 4 (dec_frame.cc:860-882): wait for the GPU, pixels are in the application's buffer
    ('  is_finalized_ = true;\n  if (decoded_->IsJPEG()) {\n    // Nothing to do.\n    return true;\n  }\n',
     '  is_finalized_ = true;\n  if (decoded_->IsJPEG()) {\n    // Nothing to do.\n    return true;\n  }\n'
     '  JXL_RETURN_IF_ERROR(jxlb_integration::FinishFrame(dec_state_));\n'),
]

DEC_GROUP = [
    ('#include "lib/jxl/dec_group.h"\n', '#include "lib/jxl/dec_group.h"\n\n' + HOOK_INCLUDE),
    # lib/jxl/dec_group.cc:335-338: the varblock's offset inside the group, for the sparse sink
    ('            qblock[c] = dec_state->coefficients->PlaneRow(c, group_idx, offset);\n          }\n',
     '            qblock[c] = dec_state->coefficients->PlaneRow(c, group_idx, offset);\n          }\n'
     '          if (jxlb_integration::SparseSink* jxlb_s = jxlb_integration::CurrentSink()) {\n'
     '            jxlb_s->offset = static_cast<uint32_t>(offset);\n          }\n'),
    # lib/jxl/dec_group.cc:483-485 + 527-531: append non-zeros to the sink instead of adding into a zero-filled block
    ('  int32_t predicted_nzeros =\n      PredictFromTopAndLeft(row_nzeros_top, row_nzeros, bx, 32);\n',
     '  jxlb_integration::SparseSink* const jxlb_sink = jxlb_integration::CurrentSink();\n'
     '  int32_t predicted_nzeros =\n      PredictFromTopAndLeft(row_nzeros_top, row_nzeros, bx, 32);\n'),
    ('    if (ac_type == ACType::k16) {\n      block.ptr16[order[k]] += coeff;\n    } else {\n'
     '      block.ptr32[order[k]] += coeff;\n    }\n',
     '    if (jxlb_sink) {\n      if (u_coeff != 0) jxlb_sink->Append(c, order[k], coeff);\n'
     '    } else if (ac_type == ACType::k16) {\n      block.ptr16[order[k]] += coeff;\n    } else {\n'
     '      block.ptr32[order[k]] += coeff;\n    }\n'),
    # lib/jxl/dec_group.cc:724-727: a GPU frame's groups are entropy-decoded only
    ('          ? kDraw\n          : kDontDraw;\n',
     '          ? kDraw\n          : kDontDraw;\n  if (jxlb_integration::DontDraw(dec_state)) draw = kDontDraw;\n'),
]


def apply(text: str, edits, name: str) -> str:
    for old, new in edits:
        n = text.count(old)
        if n != 1:
            raise SystemExit(f"patch_libjxl: anchor matches {n} times in {name} (expected 1):\n{old}")
        text = text.replace(old, new)
    return text


def main() -> int:
    ref, out = Path(sys.argv[1]), Path(sys.argv[2])
    # The copies keep their path below the output directory: lib/jxl/dec_group.cc re-includes ITSELF once per
    # Highway target (HWY_TARGET_INCLUDE + hwy/foreach_target.h), so the output directory must come first on
    # the include path -- otherwise every target but the static one would be compiled from the unpatched file.
    for rel, edits in (("lib/jxl/dec_frame.cc", DEC_FRAME), ("lib/jxl/dec_group.cc", DEC_GROUP)):
        src = (ref / rel).read_text()
        dst = out / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        new = apply(src, edits, rel)
        if not dst.exists() or dst.read_text() != new:
            dst.write_text(new)
    return 0


if __name__ == "__main__":
    sys.exit(main())
