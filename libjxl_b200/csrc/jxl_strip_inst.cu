// jxl_strip_inst.cu -- one stage chain of the row-streaming filter kernel per translation unit.
// Compiled eight times with -DSTRIP_MASK=16|17|20|21|28|29|30|31 (the chains PassesDecoderState::
// PreparePipeline can build for a VarDCT XYB frame, lib/jxl/dec_cache.cc:151-170) and linked into
// libjxl_b200.so; see libjxl_b200/pipeline.py:build().
#define JXLB_STRIP_TU 1
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "jxl_kernels.cuh"

#ifndef STRIP_MASK
#error "compile with -DSTRIP_MASK=<stage mask>"
#endif

namespace jxlb {

template <>
__attribute__((visibility("hidden"))) cudaError_t prepare_strip_mask<STRIP_MASK>() {
  constexpr uint32_t MASK = STRIP_MASK;
  cudaError_t e = cudaSuccess;
  const int bytes = (int)StripCfg<MASK>::kSmemBytes;
  auto set = [&](auto kernel) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  };
  set(filter_strip_kernel<MASK, false, 0>);
  set(filter_strip_kernel<MASK, false, 1>);
  set(filter_strip_kernel<MASK, true, 0>);
  set(filter_strip_kernel<MASK, true, 1>);
  set(filter_strip_kernel<MASK, false, 2>);
  return e;
}

template <>
cudaError_t launch_strip_mask<STRIP_MASK>(const FrameDev& P, char* dev_out, size_t out_row_bytes, int num_sms,
                                          cudaStream_t s) {
  constexpr uint32_t MASK = STRIP_MASK;
  using C = StripCfg<MASK>;
  const bool repl = P.mc || P.nrep;  // multi-GPU: the instantiation with the fused all-gather replay
  const bool plain = P.out_format == 0 && !(P.stage_mask & 32u);  // linear interleaved f32
  static int blocks_per_sm = 0;  // occupancy of this chain (all four instantiations share the bounds)
  if (!blocks_per_sm) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, filter_strip_kernel<MASK, false, 0>,
                                                                  kStripThreads, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    blocks_per_sm = n < 1 ? 1 : n;
  }
  const int band_h = (int)(P.band_y1 - P.band_y0);
  const int strips = ((int)P.xsize + C::kOutCols - 1) / C::kOutCols;
  // Exactly one wave: as many CTAs as fit on the chip at this kernel's occupancy (a 5% second
  // wave would double the kernel time), segments long enough to amortise the pipeline fill.
  const int slots = num_sms * blocks_per_sm;
  int segs = slots / strips;
  if (segs < 1) segs = 1;
  int seg_rows = (band_h + segs - 1) / segs;
  if (seg_rows < 64) seg_rows = 64;
  seg_rows = (seg_rows + 7) & ~7;
  segs = (band_h + seg_rows - 1) / seg_rows;
  const dim3 grid(strips, segs);
  static const bool special8 = [] { const char* e = getenv("JXLGPU_SRGB8_SPECIAL"); return e && e[0] == '1'; }();
  if (special8 && !repl && P.out_format == 2 && (P.stage_mask & 32u)) {  // EXPERIMENT, see store_px<2>
    filter_strip_kernel<MASK, false, 2><<<grid, kStripThreads, C::kSmemBytes, s>>>(P, dev_out, out_row_bytes, seg_rows);
    return cudaGetLastError();
  }
  if (repl) {
    if (plain) filter_strip_kernel<MASK, true, 0><<<grid, kStripThreads, C::kSmemBytes, s>>>(P, dev_out, out_row_bytes, seg_rows);
    else filter_strip_kernel<MASK, true, 1><<<grid, kStripThreads, C::kSmemBytes, s>>>(P, dev_out, out_row_bytes, seg_rows);
  } else {
    if (plain) filter_strip_kernel<MASK, false, 0><<<grid, kStripThreads, C::kSmemBytes, s>>>(P, dev_out, out_row_bytes, seg_rows);
    else filter_strip_kernel<MASK, false, 1><<<grid, kStripThreads, C::kSmemBytes, s>>>(P, dev_out, out_row_bytes, seg_rows);
  }
  return cudaGetLastError();
}

}  // namespace jxlb
