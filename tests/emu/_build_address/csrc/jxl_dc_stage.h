// jxl_dc_stage.h -- per-pixel arithmetic of the DC stage that precedes the hot path:
//   DequantDC, 4:4:4 branch      (lib/jxl/compressed_dc.cc:199-232)
//   AdaptiveDCSmoothing          (lib/jxl/compressed_dc.cc:50-197)
// Plain C++ so that the very same functions run in the CUDA kernels (dc_dequant_kernel / dc_smooth_kernel
// in jxl_kernels.cuh) and, compiled for the host by tests/test_host_logic.py, against the CPU restatement
// on a machine without a GPU.  Operation order and the places of the FMAs follow the reference; compile with
// contraction off (-fmad=false / -ffp-contract=off).
#ifndef JXL_B200_DC_STAGE_H_
#define JXL_B200_DC_STAGE_H_

#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define JXLB_HD __host__ __device__ __forceinline__
#else
#define JXLB_HD inline
#endif

namespace jxlb {

struct DcStage {
  uint32_t xb, yb;          // DC image size = frame size in 8x8 blocks
  const int32_t* q[3];      // quantised DC planes X, Y, B (dense [yb][xb])
  float* deq[3];            // dequantised planes
  float* out[3];            // smoothed planes (what the transform kernels read)
  float dc_factors[3];      // quantizer.MulDC()
  float cfl_x, cfl_b;       // cmap.base().DCFactors()[0] and [2]
  const float* group_mul;   // per DC group (256x256 blocks) 1 / (1 << extra_precision), or nullptr (= 1)
  uint32_t xdg;             // DC groups per row
};

// one block of DequantDC: in_c = float(q_c) * (dc_factors[c] * mul); Y as is; X, B += CfL * Y (one FMA)
JXLB_HD void dc_dequant_px(const DcStage& S, uint32_t x, uint32_t y) {
  const size_t i = (size_t)y * S.xb + x;
  const float mul = S.group_mul ? S.group_mul[(size_t)(y >> 8) * S.xdg + (x >> 8)] : 1.0f;
  const float fac_x = S.dc_factors[0] * mul, fac_y = S.dc_factors[1] * mul, fac_b = S.dc_factors[2] * mul;
  const float in_x = (float)S.q[0][i] * fac_x;
  const float in_y = (float)S.q[1][i] * fac_y;
  const float in_b = (float)S.q[2][i] * fac_b;
  S.deq[1][i] = in_y;
  S.deq[0][i] = fmaf(in_y, S.cfl_x, in_x);
  S.deq[2][i] = fmaf(in_y, S.cfl_b, in_b);
}

// one block of AdaptiveDCSmoothing (ComputePixel, compressed_dc.cc:96-126); border blocks are copied
JXLB_HD void dc_smooth_px(const DcStage& S, uint32_t x, uint32_t y, bool smoothing) {
  const size_t i = (size_t)y * S.xb + x;
  const bool interior = smoothing && S.xb > 2 && S.yb > 2 && x >= 1 && y >= 1 && x + 1 < S.xb && y + 1 < S.yb;
  if (!interior) {
    for (int c = 0; c < 3; c++) S.out[c][i] = S.deq[c][i];
    return;
  }
  const float w1 = 0.20345139757231578f, w2 = 0.0334829185968739f;
  const float w0 = 1.0f - 4.0f * (w1 + w2);
  const size_t up = i - S.xb, dn = i + S.xb;
  float mc[3], sm[3];
  float gap = 0.5f;
  for (int c = 0; c < 3; c++) {
    const float* p = S.deq[c];
    const float corner = (p[up - 1] + p[up + 1]) + (p[dn - 1] + p[dn + 1]);
    const float side = (p[i - 1] + p[i + 1]) + (p[up] + p[dn]);
    mc[c] = p[i];
    sm[c] = fmaf(corner, w2, fmaf(side, w1, mc[c] * w0));
    const float g = fabsf((mc[c] - sm[c]) / S.dc_factors[c]);
    gap = gap > g ? gap : g;
  }
  float factor = fmaf(-4.0f, gap, 3.0f);
  if (factor < 0.0f) factor = 0.0f;
  for (int c = 0; c < 3; c++) S.out[c][i] = fmaf(sm[c] - mc[c], factor, mc[c]);
}

}  // namespace jxlb
#endif  // JXL_B200_DC_STAGE_H_
