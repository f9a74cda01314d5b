// jxl_fused_inst.cu -- one stage chain of the fused decode kernel (jxl_fused.cuh) per translation unit.
// Compiled with -DFUSED_MASK=16|17|20|21|28|29|30 and linked into libjxl_b200.so (libjxl_b200/pipeline.py).
#define JXLB_STRIP_TU 1
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "jxl_kernels.cuh"
#include "jxl_fused.cuh"

#ifndef FUSED_MASK
#error "compile with -DFUSED_MASK=<stage mask>"
#endif

namespace jxlb {

template <>
__attribute__((visibility("hidden"))) cudaError_t prepare_fused_mask<FUSED_MASK>() {
  constexpr uint32_t MASK = FUSED_MASK;
  cudaError_t e = cudaSuccess;
  const int bytes = (int)TileCfg<MASK>::kSmemBytes;
  auto set = [&](auto kernel) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  };
  set(fused_tile_kernel<MASK, false, 0>);
  set(fused_tile_kernel<MASK, false, 1>);
  set(fused_tile_kernel<MASK, true, 0>);
  set(fused_tile_kernel<MASK, true, 1>);
  return e;
}

// Work units = strips x row segments, dealt round-robin to one persistent CTA per SM.  The segment
// count is chosen so that the units fill whole rounds of the grid (a 5 % second round would cost a
// whole round) while segments stay long enough to amortise the pipeline fill and the redundant
// halo block rows.
template <>
cudaError_t launch_fused_mask<FUSED_MASK>(const FrameDev& P, char* dev_out, size_t out_row_bytes, int num_sms,
                                          cudaStream_t s) {
  constexpr uint32_t MASK = FUSED_MASK;
  using C = TileCfg<MASK>;
  const int band_h = (int)(P.band_y1 - P.band_y0);
  const int strips = ((int)P.xsize + kTOut - 1) / kTOut;
  int best_segs = 1;
  double best = -1.0;
  for (int segs = 1; segs <= 64; segs++) {
    int seg_rows = (band_h + segs - 1) / segs;
    seg_rows = (seg_rows + 7) & ~7;
    if (segs > 1 && seg_rows < 96) break;
    const int real_segs = (band_h + seg_rows - 1) / seg_rows;
    const int units = strips * real_segs;
    const int rounds = (units + num_sms - 1) / num_sms;
    // efficiency of the rounds x overhead of short segments (fill: ~2 stages x 8 rows + 2 halo block rows)
    const double eff = (double)units / (rounds * num_sms) * seg_rows / (seg_rows + 8.0 * (C::nst + 2));
    if (eff > best) {
      best = eff;
      best_segs = real_segs;
    }
  }
  if (const char* e = getenv("JXLGPU_FUSED_SEGS")) best_segs = std::max(1, atoi(e));  // (tests: force row segments)
  int seg_rows = (band_h + best_segs - 1) / best_segs;
  seg_rows = (seg_rows + 7) & ~7;
  const int segs = (band_h + seg_rows - 1) / seg_rows;
  const int units = strips * segs;
  const int grid = std::min(units, num_sms);
  const bool plain = P.out_format == 0 && !(P.stage_mask & 32u);  // linear interleaved f32
  static const bool traced = [] { return getenv("JXLGPU_TRACE") != nullptr; }();
  if (traced)
    fprintf(stderr, "[jxl_b200] fused chain %u: %d strips x %d segments of %d rows = %d units on %d CTAs, %zu B smem\n",
            (unsigned)MASK, strips, segs, seg_rows, units, grid, (size_t)C::kSmemBytes);
  if (P.ac_is32) {
    if (plain) fused_tile_kernel<MASK, true, 0><<<grid, kTThreads, C::kSmemBytes, s>>>(P, dev_out, out_row_bytes, seg_rows, strips, units);
    else fused_tile_kernel<MASK, true, 1><<<grid, kTThreads, C::kSmemBytes, s>>>(P, dev_out, out_row_bytes, seg_rows, strips, units);
  } else {
    if (plain) fused_tile_kernel<MASK, false, 0><<<grid, kTThreads, C::kSmemBytes, s>>>(P, dev_out, out_row_bytes, seg_rows, strips, units);
    else fused_tile_kernel<MASK, false, 1><<<grid, kTThreads, C::kSmemBytes, s>>>(P, dev_out, out_row_bytes, seg_rows, strips, units);
  }
  return cudaGetLastError();
}

}  // namespace jxlb
