"""Band sharding of one frame across ranks.

AC groups are independent for dequant+IDCT; the filters need `halo` rows of post-IDCT
pixels from the neighbouring bands (Gaborish 1 + EPF 3+2+1 = 7 at most, LoopFilter::Padding,
lib/jxl/loop_filter.h:26-29).  Each rank therefore receives the coefficient groups of its own
group rows plus the adjacent group rows that intersect the halo, inverse-transforms only the
varblocks touching its rows (the plan kernel filters them) and filters its own band; the
finished bands are all-gathered.  No exchange of intermediate data between GPUs.
"""
from __future__ import annotations

from . import abi


def band_partition(ysize_groups: int, world: int) -> list[tuple[int, int]]:
    """Contiguous (y0_groups, ny_groups) per rank, sizes differ by at most one (larger first).
    Ranks beyond the number of group rows get (ysize_groups, 0)."""
    base, extra = divmod(ysize_groups, world)
    out, y = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((y, n))
        y += n
    return out


def filter_halo(desc: abi.FrameDesc) -> int:
    from .abi import STAGE_EPF0, STAGE_EPF1, STAGE_EPF2, STAGE_EXPLICIT, STAGE_GAB
    if desc.stage_mask & STAGE_EXPLICIT:
        m = desc.stage_mask
    else:
        m = (STAGE_GAB if desc.gab else 0) | (STAGE_EPF0 if desc.epf_iters >= 3 else 0) | \
            (STAGE_EPF1 if desc.epf_iters >= 1 else 0) | (STAGE_EPF2 if desc.epf_iters >= 2 else 0)
    return (1 if m & STAGE_GAB else 0) + (3 if m & STAGE_EPF0 else 0) + (2 if m & STAGE_EPF1 else 0) + \
        (1 if m & STAGE_EPF2 else 0)


def groups_needed(desc: abi.FrameDesc, y0_groups: int, ny_groups: int) -> list[int]:
    """AC groups a rank must receive to render group rows [y0, y0+ny)."""
    if ny_groups == 0:
        return []
    halo = filter_halo(desc)
    px0 = max(0, y0_groups * abi.GROUP_DIM - halo)
    px1 = min(desc.ysize, (y0_groups + ny_groups) * abi.GROUP_DIM + halo)
    gy0, gy1 = px0 // abi.GROUP_DIM, (px1 + abi.GROUP_DIM - 1) // abi.GROUP_DIM
    xg = desc.xsize_groups
    return [gy * xg + gx for gy in range(gy0, gy1) for gx in range(xg)]


def band_pixel_rows(desc: abi.FrameDesc, y0_groups: int, ny_groups: int) -> tuple[int, int]:
    y0 = min(desc.ysize, y0_groups * abi.GROUP_DIM)
    y1 = min(desc.ysize, (y0_groups + ny_groups) * abi.GROUP_DIM)
    return y0, y1 - y0
