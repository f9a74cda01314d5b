// jxl_strip2.cuh -- EXPERIMENT (branch exp/two-columns, not measured yet): the row-streaming filter
// kernel with TWO adjacent columns per thread (128 threads own the same 256-column strip).
//
// Why: the one-column kernel is FP32-issue bound; its steady step is ~352 instructions per pixel of
// which 66 are shared-memory loads and ~95 integer/control.  A thread that owns columns (2t, 2t+1)
//   * reads every ring row with three 8-byte loads (columns 2t-2 .. 2t+3) instead of 2 x 5 scalar ones,
//   * pays the per-step address arithmetic, predicates and the barrier once for two pixels,
//   * can share |a-b| terms between the two overlapping EPF diamonds.
// The strip origin is x0 - Hp with Hp = halo rounded up to even, so that x is even: a pair never
// straddles an 8x8 block (one sigma per thread) and 8-byte accesses are aligned.
// Chains with EPF0 (7x7 window) stay on the one-column kernel.
// Arithmetic (operation order, FMAs) is identical to filter_strip_body / the reference stages.
#pragma once

namespace jxlb {

constexpr int kStrip2Threads = 128;
constexpr int kStripCols = 256;

template <uint32_t MASK>
struct Strip2Cfg {
  using C = StripCfg<MASK>;
  static constexpr int Hp = (C::H + 1) & ~1;
  static constexpr int kOutCols = kStripCols - 2 * Hp;
};

template <uint32_t MASK, bool EDGE, bool REPL, int OUTK>
__device__ __forceinline__ void filter_strip2_body(const FrameDev& P, char* __restrict__ out, size_t out_row_stride,
                                                   int x0, int y_begin, int y_end, float* smem) {
  using C = StripCfg<MASK>;
  static_assert(!C::E0, "EPF0 chains use the one-column kernel");
  constexpr int H = C::H;
  constexpr int Hp = Strip2Cfg<MASK>::Hp;
  constexpr int RW = kStripCols;  // floats per ring row and channel
  const int t = threadIdx.x;
  const int c0 = 2 * t;           // strip column of pixel A; pixel B is c0 + 1
  const int W = (int)P.xsize, HI = (int)P.ysize;
  const int x = x0 - Hp + c0;     // image column of pixel A (even)
  const bool xinA = x >= 0 && x < W;
  const bool xinB = xinA && x + 1 < W;
  const int xs = min(max(x, 0), W - 1) >> 3;
  // strip-relative indices of image columns x-2 .. x+3 (mirrored at the image edge in edge strips)
  int cn[6];
#pragma unroll
  for (int k = 0; k < 6; k++) cn[k] = EDGE ? (mirror_i(x - 2 + k, W) - (x0 - Hp)) : (c0 - 2 + k);
  // Mirror() (lib/jxl/image_ops.h:184-196) reflects repeatedly: images lower than a stage's border
  auto mrow = [&](int r) { return mirror_i(r, HI); };

  float* ringG = smem + kStripPad;
  float* ring1 = ringG + C::NG * 3 * RW;
  float* ring2 = ring1 + C::N1 * 3 * RW;

  constexpr int hG = C::G ? 1 : 0;
  constexpr int h1 = hG + (C::E1 ? 2 : 0);
  constexpr int h2 = h1 + (C::E2 ? 1 : 0);
  static_assert(h2 == H, "halo bookkeeping");
  auto lo = [&](int rem) { return max(0, y_begin - rem); };
  auto hi = [&](int rem) { return min(HI, y_end + rem); };

  const float kMinSigma = -3.90524291751269967465540850526868f;
  const bool xborderA = (x & 7) == 0, xborderB = (x & 7) == 6;   // x even: A sits on 0/2/4/6, B on 1/3/5/7
  const int band_h = (int)P.out_h;
  const bool emitA = xinA && c0 >= Hp && c0 < kStripCols - Hp;
  const bool emitB = emitA && xinB;
  // edge strips: lanes outside the image would mirror to columns the strip does not hold
  const bool lane_ok = EDGE ? xinA : true;

  auto xyb_to_rgb = [&](float& a, float& b, float& c3) {
    if constexpr (C::XYB) {
      float gr = b + a, gg = b - a, gb = c3;
      gr = gr - P.opsin_cbrt[0];
      gg = gg - P.opsin_cbrt[1];
      gb = gb - P.opsin_cbrt[2];
      const float r2 = gr * gr, g2 = gg * gg, b2 = gb * gb;
      const float mr = fmaf(r2, gr, P.opsin_bias[0]);
      const float mg = fmaf(g2, gg, P.opsin_bias[1]);
      const float mb = fmaf(b2, gb, P.opsin_bias[2]);
      float lr = P.opsin_m[0] * mr, lg = P.opsin_m[3] * mr, lb = P.opsin_m[6] * mr;
      lr = fmaf(P.opsin_m[1], mg, lr); lg = fmaf(P.opsin_m[4], mg, lg); lb = fmaf(P.opsin_m[7], mg, lb);
      lr = fmaf(P.opsin_m[2], mb, lr); lg = fmaf(P.opsin_m[5], mb, lg); lb = fmaf(P.opsin_m[8], mb, lb);
      a = lr; b = lg; c3 = lb;
    }
  };
  // pa / pb: X, Y, B of the two pixels of row r
  auto emit2 = [&](int r, const float* pa, const float* pb) {
    if (!emitA) return;
    float a0 = pa[0], a1 = pa[1], a2 = pa[2];
    xyb_to_rgb(a0, a1, a2);
    store_px<OUTK>(P, out, out_row_stride, r - (int)P.out_y0, x, band_h, a0, a1, a2);
    if (emitB) {
      float b0 = pb[0], b1 = pb[1], b2 = pb[2];
      xyb_to_rgb(b0, b1, b2);
      store_px<OUTK>(P, out, out_row_stride, r - (int)P.out_y0, x + 1, band_h, b0, b1, b2);
    }
  };

  constexpr int dG = C::G ? 2 : 0;
  constexpr int d1 = dG + (C::E1 ? 3 : 0);
  constexpr int d2 = d1 + (C::E2 ? 2 : 0);
  const int r_in_lo = lo(H), r_in_hi = hi(H);
  const int r_end = hi(0) + d2;

  // row r_in_lo is fetched up front, every later row one step ahead of its use
  float preA[3] = {0.0f, 0.0f, 0.0f}, preB[3] = {0.0f, 0.0f, 0.0f};
  auto fetch = [&](int row) {
    const size_t off = (size_t)row * P.row_stride + x;
    if (xinB) {  // x is even, rows and planes start on 32-byte boundaries: an aligned 8-byte load
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(P.xyb + c * P.plane_stride + off));
        preA[c] = v.x;
        preB[c] = v.y;
      }
    } else if (xinA) {
#pragma unroll
      for (int c = 0; c < 3; c++) preA[c] = __ldg(P.xyb + c * P.plane_stride + off);
    }
  };
  if (r_in_lo < r_in_hi) fetch(r_in_lo);

  float sg1 = 0.0f, sg2 = 0.0f;
  if (C::E1) sg1 = __ldg(P.sigma + (size_t)(lo(H - h1) >> 3) * P.xb + xs);
  if (C::E2) sg2 = __ldg(P.sigma + (size_t)(lo(0) >> 3) * P.xb + xs);

  // row loads: own pair, the pair plus one column each side, plus two columns each side
  auto LD2 = [&](const float* rowc, float* v) {
    const float2 a = *reinterpret_cast<const float2*>(rowc + c0);
    v[0] = a.x; v[1] = a.y;
  };
  auto LD4 = [&](const float* rowc, float* v) {  // columns c0-1 .. c0+2
    if constexpr (EDGE) {
      v[0] = rowc[cn[1]]; v[1] = rowc[cn[2]]; v[2] = rowc[cn[3]]; v[3] = rowc[cn[4]];
    } else {
      const float2* p = reinterpret_cast<const float2*>(rowc + c0);
      const float2 a = p[-1], b = p[0], d = p[1];
      v[0] = a.y; v[1] = b.x; v[2] = b.y; v[3] = d.x;
    }
  };
  auto LD6 = [&](const float* rowc, float* v) {  // columns c0-2 .. c0+3
    if constexpr (EDGE) {
#pragma unroll
      for (int k = 0; k < 6; k++) v[k] = rowc[cn[k]];
    } else {
      const float2* p = reinterpret_cast<const float2*>(rowc + c0);
      const float2 a = p[-1], b = p[0], d = p[1];
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = d.x; v[5] = d.y;
    }
  };
  auto ST2 = [&](float* rowc, float a, float b) { *reinterpret_cast<float2*>(rowc + c0) = make_float2(a, b); };

  // J >= 0 (aligned): rin == 8*m + J, so every ring slot is a compile-time constant (as in filter_strip_body)
  auto step = [&](auto steady_tag, auto jtag, int rin) {
    constexpr bool ST = decltype(steady_tag)::value;
    constexpr int J = decltype(jtag)::value;
    static_assert(J < 0 || ST, "aligned steps are steady steps");
    auto mr = [&](int r) { return ST ? r : mrow(r); };
    // channel-0 row pointer of ring row (rin + dk); r_dyn is that row (mirrored in generic mode)
    auto RP = [&](float* ring, auto ntag, auto dktag, int r_dyn) -> float* {
      constexpr int n = decltype(ntag)::value;
      constexpr int dk = decltype(dktag)::value;
      if constexpr (J >= 0) return ring + ((((J + dk) % n + n) % n) * 3) * RW;
      else return ring + ((r_dyn & (n - 1)) * 3) * RW;
    };
    // which: 0 = loader output, 1 = Gaborish, 3 = EPF1, 4 = EPF2
    auto deliver2 = [&](auto which_tag, int r, const float* pa, const float* pb) {
      constexpr int which = decltype(which_tag)::value;
      constexpr int D = which == 0 ? 0 : (which == 1 ? dG : (which == 3 ? d1 : d2));
      constexpr bool toG = which < 1 && C::G;
      constexpr bool to1 = !toG && which < 3 && C::E1;
      constexpr bool to2 = !toG && !to1 && which < 4 && C::E2;
      float* dst = nullptr;
      if constexpr (toG) dst = RP(ringG, IC<C::NG ? C::NG : 1>(), IC<-D>(), r);
      else if constexpr (to1) dst = RP(ring1, IC<C::N1 ? C::N1 : 1>(), IC<-D>(), r);
      else if constexpr (to2) dst = RP(ring2, IC<C::N2 ? C::N2 : 1>(), IC<-D>(), r);
      if constexpr (toG || to1 || to2) {
        ST2(dst, pa[0], pb[0]);
        ST2(dst + RW, pa[1], pb[1]);
        ST2(dst + 2 * RW, pa[2], pb[2]);
      } else {
        emit2(r, pa, pb);
      }
    };
    // ---- loader ----
    if ((ST || rin < r_in_hi) && xinA) deliver2(IC<0>(), rin, preA, preB);
    if (rin + 1 < r_in_hi) fetch(rin + 1);
    // ---- Gaborish (stage_gaborish.cc:56-100) ----
    if constexpr (C::G) {
      const int r = rin - dG;
      if ((ST || (r >= lo(H - hG) && r < hi(H - hG))) && lane_ok) {
        const float* pT0 = RP(ringG, IC<C::NG>(), IC<-dG - 1>(), mr(r - 1));
        const float* pM0 = RP(ringG, IC<C::NG>(), IC<-dG>(), r);
        const float* pB0 = RP(ringG, IC<C::NG>(), IC<-dG + 1>(), mr(r + 1));
        float oa[3], ob[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          float T[4], M[4], B[4];
          LD4(pT0 + c * RW, T);
          LD4(pM0 + c * RW, M);
          LD4(pB0 + c * RW, B);
          const float w0 = P.gab_w[3 * c], w1 = P.gab_w[3 * c + 1], w2 = P.gab_w[3 * c + 2];
          {
            const float sum1 = (M[0] + M[2]) + (T[1] + B[1]);
            const float sum2 = (T[0] + T[2]) + (B[0] + B[2]);
            oa[c] = fmaf(sum2, w2, fmaf(sum1, w1, M[1] * w0));
          }
          {
            const float sum1 = (M[1] + M[3]) + (T[2] + B[2]);
            const float sum2 = (T[1] + T[3]) + (B[1] + B[3]);
            ob[c] = fmaf(sum2, w2, fmaf(sum1, w1, M[2] * w0));
          }
        }
        deliver2(IC<1>(), r, oa, ob);
      }
    }
    // ---- EPF1 (stage_epf.cc:197-379) ----
    if constexpr (C::E1) {
      const int r = rin - d1;
      if ((ST || (r >= lo(H - h1) && r < hi(H - h1))) && lane_ok) {
        const float s = sg1;
        sg1 = __ldg(P.sigma + (size_t)(min(max(r + 1, 0), HI - 1) >> 3) * P.xb + xs);
        const float* q2x = RP(ring1, IC<C::N1>(), IC<-d1>(), r);
        float pa[3], pb[3];
        if (!(s < kMinSigma)) {
          const int iy = r & 7;
          const float sm_ = P.epf_sm[1];
          const bool yb_ = iy == 0 || iy == 7;
          const float inv_sigma_a = s * ((yb_ || xborderA) ? sm_ * P.epf_border_mul : sm_);
          const float inv_sigma_b = s * ((yb_ || xborderB) ? sm_ * P.epf_border_mul : sm_);
          const float* q0x = RP(ring1, IC<C::N1>(), IC<-d1 - 2>(), mr(r - 2));
          const float* q1x = RP(ring1, IC<C::N1>(), IC<-d1 - 1>(), mr(r - 1));
          const float* q3x = RP(ring1, IC<C::N1>(), IC<-d1 + 1>(), mr(r + 1));
          const float* q4x = RP(ring1, IC<C::N1>(), IC<-d1 + 2>(), mr(r + 2));
          float sada[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sadb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          float nba[3][4], nbb[3][4];  // neighbour pixels N, W, E, S per channel
          float ca[3], cb[3];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            float r0[2], r1[4], r2[6], r3[4], r4[2];
            LD2(q0x + c * RW, r0);
            LD4(q1x + c * RW, r1);
            LD6(q2x + c * RW, r2);
            LD4(q3x + c * RW, r3);
            LD2(q4x + c * RW, r4);
            const float scale = P.epf_scale[c];
            // |a-b| terms of pixel A (named after the order filter_strip_body uses them in) ...
            const float A1 = fabsf(r0[0] - r1[1]), A2 = fabsf(r1[0] - r1[1]), A3 = fabsf(r1[2] - r1[1]);
            const float A4 = fabsf(r2[0] - r2[1]), A5 = fabsf(r1[0] - r2[1]), A6 = fabsf(r2[1] - r2[2]);
            const float A7 = fabsf(r2[2] - r1[1]), A8 = fabsf(r1[2] - r2[3]), A9 = fabsf(r2[2] - r2[3]);
            const float A10 = fabsf(r2[4] - r2[3]), A11 = fabsf(r3[0] - r2[1]), A12 = fabsf(r2[2] - r3[1]);
            const float A13 = fabsf(r3[0] - r3[1]), A14 = fabsf(r3[2] - r3[1]), A15 = fabsf(r3[2] - r2[3]);
            const float A16 = fabsf(r4[0] - r3[1]);
            // ... and of pixel B, one column to the right: 9 of its 16 pairs are pairs A already has
            // (|x - y| == |y - x| exactly), 7 are new
            const float B1 = fabsf(r0[1] - r1[2]), B2 = A3, B3 = fabsf(r1[3] - r1[2]), B4 = A6, B5 = A7, B6 = A9, B7 = A8;
            const float B8 = fabsf(r1[3] - r2[4]), B9 = A10, B10 = fabsf(r2[5] - r2[4]), B11 = A12, B12 = A15, B13 = A14;
            const float B14 = fabsf(r3[3] - r3[2]), B15 = fabsf(r3[3] - r2[4]), B16 = fabsf(r4[1] - r3[2]);
            // accumulation in the reference's order (stage_epf.cc:197-379)
            auto acc = [&](float t1, float t2, float t3, float t4, float t5, float t6, float t7, float t8, float t9,
                           float t10, float t11, float t12, float t13, float t14, float t15, float t16, float* sad) {
              float sad0c = t1;
              float sad1c = t2;
              float sad2c = t3;
              sad1c = sad1c + t4;
              sad0c = sad0c + t5;
              sad1c = sad1c + t6;
              sad2c = sad2c + t6;
              float sad3c = t7;
              sad0c = sad0c + t7;
              sad0c = sad0c + t8;
              sad1c = sad1c + t9;
              sad2c = sad2c + t9;
              sad2c = sad2c + t10;
              sad3c = sad3c + t11;
              sad0c = sad0c + t12;
              sad3c = sad3c + t12;
              sad1c = sad1c + t13;
              sad2c = sad2c + t14;
              sad3c = sad3c + t15;
              sad3c = sad3c + t16;
              sad[0] = fmaf(sad0c, scale, sad[0]);
              sad[1] = fmaf(sad1c, scale, sad[1]);
              sad[2] = fmaf(sad2c, scale, sad[2]);
              sad[3] = fmaf(sad3c, scale, sad[3]);
            };
            acc(A1, A2, A3, A4, A5, A6, A7, A8, A9, A10, A11, A12, A13, A14, A15, A16, sada);
            acc(B1, B2, B3, B4, B5, B6, B7, B8, B9, B10, B11, B12, B13, B14, B15, B16, sadb);
            // neighbours N, W, E, S and the centre of each pixel
            nba[c][0] = r1[1]; nba[c][1] = r2[1]; nba[c][2] = r2[3]; nba[c][3] = r3[1]; ca[c] = r2[2];
            nbb[c][0] = r1[2]; nbb[c][1] = r2[2]; nbb[c][2] = r2[4]; nbb[c][3] = r3[2]; cb[c] = r2[3];
          }
          auto finish = [&](const float* sad, float (*nb)[4], const float* ctr, float inv_sigma, float* o) {
            float X = ctr[0], Y = ctr[1], B = ctr[2];
            float w = 1.0f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const float wt = epf_weight(sad[k], inv_sigma);
              w = w + wt;
              X = fmaf(wt, nb[0][k], X);
              Y = fmaf(wt, nb[1][k], Y);
              B = fmaf(wt, nb[2][k], B);
            }
            const float inv_w = 1.0f / w;
            o[0] = X * inv_w; o[1] = Y * inv_w; o[2] = B * inv_w;
          };
          finish(sada, nba, ca, inv_sigma_a, pa);
          finish(sadb, nbb, cb, inv_sigma_b, pb);
        } else {
#pragma unroll
          for (int c = 0; c < 3; c++) {
            float v[2];
            LD2(q2x + c * RW, v);
            pa[c] = v[0];
            pb[c] = v[1];
          }
        }
        deliver2(IC<3>(), r, pa, pb);
      }
    }
    // ---- EPF2 (stage_epf.cc:383-506) ----
    if constexpr (C::E2) {
      const int r = rin - d2;
      if ((ST || (r >= lo(0) && r < hi(0))) && lane_ok) {
        const float s = sg2;
        sg2 = __ldg(P.sigma + (size_t)(min(max(r + 1, 0), HI - 1) >> 3) * P.xb + xs);
        const float* pM = RP(ring2, IC<C::N2>(), IC<-d2>(), r);
        float pa[3], pb[3];
        if (!(s < kMinSigma)) {
          const int iy = r & 7;
          const float sm_ = P.epf_sm[2];
          const bool yb_ = iy == 0 || iy == 7;
          const float inv_sigma_a = s * ((yb_ || xborderA) ? sm_ * P.epf_border_mul : sm_);
          const float inv_sigma_b = s * ((yb_ || xborderB) ? sm_ * P.epf_border_mul : sm_);
          const float* pT = RP(ring2, IC<C::N2>(), IC<-d2 - 1>(), mr(r - 1));
          const float* pB = RP(ring2, IC<C::N2>(), IC<-d2 + 1>(), mr(r + 1));
          float T[3][2], M[3][4], Bt[3][2];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            LD2(pT + c * RW, T[c]);
            LD4(pM + c * RW, M[c]);
            LD2(pB + c * RW, Bt[c]);
          }
          // neighbours in the order N, W, E, S; i = 0 for pixel A, 1 for pixel B
          auto one = [&](int i, float inv_sigma, float* o) {
            const float rx = M[0][1 + i], ry = M[1][1 + i], rb = M[2][1 + i];
            float X = rx, Y = ry, B = rb;
            float w = 1.0f;
            const float nx[4] = {T[0][i], M[0][i], M[0][2 + i], Bt[0][i]};
            const float ny[4] = {T[1][i], M[1][i], M[1][2 + i], Bt[1][i]};
            const float nb[4] = {T[2][i], M[2][i], M[2][2 + i], Bt[2][i]};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              float sad = fabsf(nx[k] - rx) * P.epf_scale[0];
              sad = fmaf(fabsf(ny[k] - ry), P.epf_scale[1], sad);
              sad = fmaf(fabsf(nb[k] - rb), P.epf_scale[2], sad);
              const float wt = epf_weight(sad, inv_sigma);
              w = w + wt;
              X = fmaf(wt, nx[k], X);
              Y = fmaf(wt, ny[k], Y);
              B = fmaf(wt, nb[k], B);
            }
            const float inv_w = 1.0f / w;
            o[0] = X * inv_w; o[1] = Y * inv_w; o[2] = B * inv_w;
          };
          one(0, inv_sigma_a, pa);
          one(1, inv_sigma_b, pb);
        } else {
#pragma unroll
          for (int c = 0; c < 3; c++) {
            float v[2];
            LD2(pM + c * RW, v);
            pa[c] = v[0];
            pb[c] = v[1];
          }
        }
        deliver2(IC<4>(), r, pa, pb);
      }
    }
    if constexpr (H > 0) __syncthreads();
  };

  int s_lo = r_in_lo, s_hi = r_in_hi;
  auto constrain = [&](int d, int b, int rem) {
    s_lo = max(s_lo, max(lo(rem), b) + d);
    s_hi = min(s_hi, min(hi(rem), HI - b) + d);
  };
  if (C::G) constrain(dG, 1, H - hG);
  if (C::E1) constrain(d1, 2, H - h1);
  if (C::E2) constrain(d2, 1, 0);
  if (s_hi < s_lo) s_hi = s_lo;
  const int s_begin = min(s_lo, r_end), s_end = min(s_hi, r_end);
  using F = std::false_type;
  using T = std::true_type;
  constexpr bool replicate = REPL;
  const int ncols_out = min(Strip2Cfg<MASK>::kOutCols, W - x0);
  int replayed = y_begin;
  auto replay_to = [&](int row_excl) {
    if constexpr (H == 0) __syncthreads();
    row_excl = min(row_excl, y_end);
    for (int y = replayed; y < row_excl; y++) {
      const size_t yo = (size_t)(y - (int)P.out_y0);
      if (P.out_format == 1) {
        for (int c = 0; c < 3; c++)
          replicate_span(P, out, ((size_t)c * band_h + yo) * out_row_stride + (size_t)x0 * 4, ncols_out * 4);
      } else {
        const int pxb = out_pixel_bytes(P.out_format);
        replicate_span(P, out, yo * out_row_stride + (size_t)x0 * pxb, ncols_out * pxb);
      }
    }
    if (row_excl > replayed) replayed = row_excl;
  };
  int rin = r_in_lo;
  for (; rin < s_begin; rin++) step(F(), IC<-1>(), rin);
  if constexpr (H > 0) {
    for (; rin < s_end && (rin & 7); rin++) step(T(), IC<-1>(), rin);
    for (; rin + 8 <= s_end; rin += 8) {
      step(T(), IC<0>(), rin);
      step(T(), IC<1>(), rin + 1);
      step(T(), IC<2>(), rin + 2);
      step(T(), IC<3>(), rin + 3);
      step(T(), IC<4>(), rin + 4);
      step(T(), IC<5>(), rin + 5);
      step(T(), IC<6>(), rin + 6);
      step(T(), IC<7>(), rin + 7);
      if constexpr (replicate) replay_to(rin + 8 - d2);
    }
  }
  for (; rin < s_end; rin++) {
    step(T(), IC<-1>(), rin);
    if constexpr (replicate) {
      if ((rin & 7) == 7) replay_to(rin + 1 - d2);
    }
  }
  for (; rin < r_end; rin++) step(F(), IC<-1>(), rin);
  if constexpr (replicate) replay_to(y_end);
}

template <uint32_t MASK, bool REPL, int OUTK>
__global__ void __launch_bounds__(kStrip2Threads) filter_strip2_kernel(const __grid_constant__ FrameDev P,
                                                                       char* __restrict__ out, size_t out_row_stride,
                                                                       int seg_rows) {
  extern __shared__ __align__(16) float fsm[];
  using C2 = Strip2Cfg<MASK>;
  const int x0 = blockIdx.x * C2::kOutCols;
  const int y_begin = (int)P.band_y0 + blockIdx.y * seg_rows;
  const int y_end = min((int)P.band_y1, y_begin + seg_rows);
  if (y_begin >= y_end) return;
  const bool edge = (x0 - C2::Hp < 0) || (x0 - C2::Hp + kStripCols > (int)P.xsize);
  if (edge) filter_strip2_body<MASK, true, REPL, OUTK>(P, out, out_row_stride, x0, y_begin, y_end, fsm);
  else filter_strip2_body<MASK, false, REPL, OUTK>(P, out, out_row_stride, x0, y_begin, y_end, fsm);
}

}  // namespace jxlb
