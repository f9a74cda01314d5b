/* jxl_oracle.c -- CPU restatement of the reference's VarDCT decode transform path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library
 * (libjxl_b200/csrc) never links, loads or calls it and fails loudly without CUDA.
 *
 * PARITY IS PINNED: tests/test_oracle_vs_reference.py checks every function here
 * against the unmodified reference built by oracle/build_ref.py (all 27
 * AcStrategy transforms incl. 128/256 via jxl::TransformToPixels, LLF-from-DC,
 * and whole frames stage by stage via DecodeGroupForRoundtrip + the reference's
 * own Gaborish/EPF/XYB stages), and tests/golden/ holds committed vectors made
 * by tests/golden/make_golden.py from that same reference build.
 *
 * Plain scalar C, one operation per reference operation, fmaf() exactly where the
 * reference uses MulAdd/NegMulAdd (AVX2 target => fused).  Compile with
 * -ffp-contract=off so the compiler adds no fusions of its own.
 *
 * Each function cites the reference file:line it follows (paths under
 * /root/reference).
 */
#include "jxl_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <xmmintrin.h>
#endif

#define JXT_CONST static const
#include "../libjxl_b200/csrc/jxl_tables.h"

static const float kSqrt2 = 1.41421356237f; /* lib/jxl/dct_scales.h:15 */

static inline const float* WC(int n) { return JXT_WC + (n / 2 - 2); }
static inline const float* RESAMPLE(int n) { return JXT_RESAMPLE + (n - 1); }

/* ------------------------------------------------------------------------- */
/* 1-D transforms: lib/jxl/dct-inl.h:45-232 (Perera-Liu radix-2 recursion)    */
/* ------------------------------------------------------------------------- */

/* IDCT1DImpl<N> (dct-inl.h:191-232). tmp needs 2N floats. in/out may alias when
 * strides are equal (the reference calls itself in place, dct-inl.h:226-229). */
static void idct1d(int N, const float* in, size_t is, float* out, size_t os, float* tmp) {
  if (N == 1) { out[0] = in[0]; return; }
  if (N == 2) {
    float a = in[0], b = in[is];
    out[0] = a + b;
    out[os] = a - b;
    return;
  }
  const int H = N / 2;
  /* ForwardEvenOdd (dct-inl.h:117-128) */
  for (int i = 0; i < H; i++) tmp[i] = in[(size_t)(2 * i) * is];
  for (int i = 0; i < H; i++) tmp[H + i] = in[(size_t)(2 * i + 1) * is];
  idct1d(H, tmp, 1, tmp, 1, tmp + N);
  /* BTranspose (dct-inl.h:91-101) */
  float* o = tmp + H;
  for (int i = H - 1; i > 0; i--) o[i] = o[i] + o[i - 1];
  o[0] = o[0] * kSqrt2;
  idct1d(H, o, 1, o, 1, tmp + N);
  /* MultiplyAndAdd (dct-inl.h:139-152) */
  const float* w = WC(N);
  for (int i = 0; i < H; i++) {
    float mul = w[i], in1 = tmp[i], in2 = o[i];
    out[(size_t)i * os] = fmaf(mul, in2, in1);
    out[(size_t)(N - 1 - i) * os] = fmaf(-mul, in2, in1);
  }
}

/* DCT1DImpl<N> (dct-inl.h:158-189), in place on contiguous mem[N]; tmp 2N floats.
 * The 1/N scale of StoreToBlockAndScale (dct-inl.h:162-170) is applied by callers. */
static void dct1d(int N, float* mem, float* tmp) {
  if (N == 1) return;
  if (N == 2) {
    float a = mem[0], b = mem[1];
    mem[0] = a + b;
    mem[1] = a - b;
    return;
  }
  const int H = N / 2;
  for (int i = 0; i < H; i++) tmp[i] = mem[i] + mem[N - 1 - i];       /* AddReverse */
  dct1d(H, tmp, tmp + N);
  for (int i = 0; i < H; i++) tmp[H + i] = mem[i] - mem[N - 1 - i];   /* SubReverse */
  const float* w = WC(N);
  for (int i = 0; i < H; i++) tmp[H + i] = tmp[H + i] * w[i];         /* Multiply */
  dct1d(H, tmp + H, tmp + N);
  float* c = tmp + H;                                                 /* B (dct-inl.h:80-90) */
  c[0] = fmaf(c[0], kSqrt2, c[1]);
  for (int i = 1; i + 1 < H; i++) c[i] = c[i] + c[i + 1];
  for (int i = 0; i < H; i++) {                                       /* InverseEvenOdd */
    mem[2 * i] = tmp[i];
    mem[2 * i + 1] = tmp[H + i];
  }
}

/* ComputeScaledIDCT<ROWS,COLS> (dct-inl.h:376-397).  `from` holds ROWS*COLS
 * coefficients in layout [min(R,C)][max(R,C)] and is clobbered; scratch needs
 * ROWS*COLS + 2*max floats. */
static void scaled_idct(int R, int C, float* from, float* to, size_t to_stride, float* scratch) {
  float* block = scratch;
  float* tmp = scratch + (size_t)R * C;
  if (R < C) {
    /* Transpose [R][C] -> block [C][R] */
    for (int y = 0; y < R; y++)
      for (int x = 0; x < C; x++) block[(size_t)x * R + y] = from[(size_t)y * C + x];
    /* IDCT1D<COLS,ROWS>: N = C down the first axis, R lanes: block -> from [C][R] */
    for (int j = 0; j < R; j++) idct1d(C, block + j, R, from + j, R, tmp);
    /* Transpose [C][R] -> block [R][C] */
    for (int x = 0; x < C; x++)
      for (int y = 0; y < R; y++) block[(size_t)y * C + x] = from[(size_t)x * R + y];
    /* IDCT1D<ROWS,COLS>: N = R down the first axis, C lanes -> pixels */
    for (int x = 0; x < C; x++) idct1d(R, block + x, C, to + x, to_stride, tmp);
  } else {
    /* from is [C][R] */
    for (int j = 0; j < R; j++) idct1d(C, from + j, R, block + j, R, tmp);
    for (int x = 0; x < C; x++)
      for (int y = 0; y < R; y++) from[(size_t)y * C + x] = block[(size_t)x * R + y];
    for (int x = 0; x < C; x++) idct1d(R, from + x, C, to + x, to_stride, tmp);
  }
}

/* ComputeScaledDCT<ROWS,COLS> (dct-inl.h:349-371). from: pixels [R][C] with stride;
 * to: R*C coefficients, layout [min][max]; scratch R*C + 2*max floats. */
static void scaled_dct(int R, int C, const float* from, size_t from_stride, float* to,
                       float* scratch) {
  float* block = scratch;
  float* tmp = scratch + (size_t)R * C;
  float col[256];
  /* DCT1D<ROWS,COLS>: N = R down columns, result [R][C], scaled 1/R */
  float* first = (R < C) ? block : to;
  const float mulR = 1.0f / R, mulC = 1.0f / C;
  for (int x = 0; x < C; x++) {
    for (int y = 0; y < R; y++) col[y] = from[(size_t)y * from_stride + x];
    dct1d(R, col, tmp);
    for (int y = 0; y < R; y++) first[(size_t)y * C + x] = mulR * col[y];
  }
  if (R < C) {
    /* Transpose block [R][C] -> to [C][R]; DCT N = C down first axis -> block [C][R];
     * transpose -> to [R][C] */
    for (int y = 0; y < R; y++)
      for (int x = 0; x < C; x++) to[(size_t)x * R + y] = block[(size_t)y * C + x];
    for (int j = 0; j < R; j++) {
      for (int x = 0; x < C; x++) col[x] = to[(size_t)x * R + j];
      dct1d(C, col, tmp);
      for (int x = 0; x < C; x++) block[(size_t)x * R + j] = mulC * col[x];
    }
    for (int x = 0; x < C; x++)
      for (int y = 0; y < R; y++) to[(size_t)y * C + x] = block[(size_t)x * R + y];
  } else {
    /* Transpose to [R][C] -> block [C][R]; DCT N = C down first axis -> to [C][R] */
    for (int y = 0; y < R; y++)
      for (int x = 0; x < C; x++) block[(size_t)x * R + y] = to[(size_t)y * C + x];
    for (int j = 0; j < R; j++) {
      for (int x = 0; x < C; x++) col[x] = block[(size_t)x * R + j];
      dct1d(C, col, tmp);
      for (int x = 0; x < C; x++) to[(size_t)x * R + j] = mulC * col[x];
    }
  }
}

/* ------------------------------------------------------------------------- */
/* 8x8 specials: lib/jxl/dec_transforms-inl.h:66-93, 95-454, 463-581           */
/* ------------------------------------------------------------------------- */
static void idct2_top_block(int S, const float* block, float* out) {
  /* IDCT2TopBlock<S> (dec_transforms-inl.h:66-93), stride_out == 8 */
  float temp[64];
  const int n = S / 2;
  for (int y = 0; y < n; y++)
    for (int x = 0; x < n; x++) {
      float c00 = block[y * 8 + x];
      float c01 = block[y * 8 + n + x];
      float c10 = block[(y + n) * 8 + x];
      float c11 = block[(y + n) * 8 + n + x];
      float r00 = c00 + c01 + c10 + c11;
      float r01 = c00 + c01 - c10 - c11;
      float r10 = c00 - c01 + c10 - c11;
      float r11 = c00 - c01 - c10 + c11;
      temp[y * 2 * 8 + x * 2] = r00;
      temp[y * 2 * 8 + x * 2 + 1] = r01;
      temp[(y * 2 + 1) * 8 + x * 2] = r10;
      temp[(y * 2 + 1) * 8 + x * 2 + 1] = r11;
    }
  for (int y = 0; y < S; y++)
    for (int x = 0; x < S; x++) out[y * 8 + x] = temp[y * 8 + x];
}

static void afv_idct4x4(const float* coeffs, float* pixels) {
  /* AFVIDCT4x4 (dec_transforms-inl.h:384-397): fma chain over j ascending */
  for (int i = 0; i < 16; i++) {
    float p = 0.0f;
    for (int j = 0; j < 16; j++) p = fmaf(coeffs[j], JXT_AFV_BASIS[j][i], p);
    pixels[i] = p;
  }
}

static void afv_to_pixels(int afv_kind, const float* co, float* px, size_t stride) {
  /* AFVTransformToPixels<afv_kind> (dec_transforms-inl.h:399-454) */
  float scratch[4 * 8 * 4 + 64];
  const int afv_x = afv_kind & 1, afv_y = afv_kind / 2;
  float block00 = co[0], block01 = co[1], block10 = co[8];
  float dcs0 = (block00 + block10 + block01) * 4.0f;
  float dcs1 = (block00 + block10 - block01);
  float dcs2 = block00 - block10;
  float coeff[16], block[32];
  coeff[0] = dcs0;
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++) {
      if (ix == 0 && iy == 0) continue;
      coeff[iy * 4 + ix] = co[iy * 2 * 8 + ix * 2];
    }
  afv_idct4x4(coeff, block);
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++)
      px[(size_t)(iy + afv_y * 4) * stride + afv_x * 4 + ix] =
          block[(afv_y == 1 ? 3 - iy : iy) * 4 + (afv_x == 1 ? 3 - ix : ix)];
  block[0] = dcs1;
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++) {
      if (ix == 0 && iy == 0) continue;
      block[iy * 4 + ix] = co[iy * 2 * 8 + ix * 2 + 1];
    }
  scaled_idct(4, 4, block, px + (size_t)afv_y * 4 * stride + (afv_x == 1 ? 0 : 4), stride, scratch);
  block[0] = dcs2;
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 8; ix++) {
      if (ix == 0 && iy == 0) continue;
      block[iy * 8 + ix] = co[(1 + iy * 2) * 8 + ix];
    }
  scaled_idct(4, 8, block, px + (size_t)(afv_y == 1 ? 0 : 4) * stride, stride, scratch);
}

static const unsigned char kRows8[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1,
                                         1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16};
static const unsigned char kCols8[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1,
                                         1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32};

int jxo_covered_blocks_x(int strategy) { return kCols8[strategy]; }
int jxo_covered_blocks_y(int strategy) { return kRows8[strategy]; }

static int is_plain_dct(int s) { return s == 0 || (s >= 4 && s <= 11) || s >= 18; }

/* TransformToPixels (dec_transforms-inl.h:456-689). coefficients are clobbered.
 * scratch: 2 * R*C + 2*256 floats. */
static void transform_to_pixels(int strategy, float* co, float* px, size_t stride, float* scratch) {
  if (is_plain_dct(strategy)) {
    scaled_idct(kRows8[strategy] * 8, kCols8[strategy] * 8, co, px, stride, scratch);
    return;
  }
  switch (strategy) {
    case 1: { /* IDENTITY (463-499) */
      float b00 = co[0], b01 = co[1], b10 = co[8], b11 = co[9];
      float dcs[4];
      dcs[0] = b00 + b01 + b10 + b11;
      dcs[1] = b00 + b01 - b10 - b11;
      dcs[2] = b00 - b01 + b10 - b11;
      dcs[3] = b00 - b01 - b10 + b11;
      for (int y = 0; y < 2; y++)
        for (int x = 0; x < 2; x++) {
          float block_dc = dcs[y * 2 + x];
          float residual_sum = 0;
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 0 && iy == 0) continue;
              residual_sum += co[(y + iy * 2) * 8 + x + ix * 2];
            }
          px[(size_t)(4 * y + 1) * stride + 4 * x + 1] = block_dc - residual_sum * (1.0f / 16);
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 1 && iy == 1) continue;
              px[(size_t)(y * 4 + iy) * stride + x * 4 + ix] =
                  co[(y + iy * 2) * 8 + x + ix * 2] + px[(size_t)(4 * y + 1) * stride + 4 * x + 1];
            }
          px[(size_t)y * 4 * stride + x * 4] =
              co[(y + 2) * 8 + x + 2] + px[(size_t)(4 * y + 1) * stride + 4 * x + 1];
        }
      break;
    }
    case 13: { /* DCT8X4 (500-519) */
      float b0 = co[0], b1 = co[8];
      float dcs[2] = {b0 + b1, b0 - b1};
      for (int x = 0; x < 2; x++) {
        float block[32];
        block[0] = dcs[x];
        for (int iy = 0; iy < 4; iy++)
          for (int ix = 0; ix < 8; ix++) {
            if (ix == 0 && iy == 0) continue;
            block[iy * 8 + ix] = co[(x + iy * 2) * 8 + ix];
          }
        scaled_idct(8, 4, block, px + x * 4, stride, scratch);
      }
      break;
    }
    case 12: { /* DCT4X8 (520-540) */
      float b0 = co[0], b1 = co[8];
      float dcs[2] = {b0 + b1, b0 - b1};
      for (int y = 0; y < 2; y++) {
        float block[32];
        block[0] = dcs[y];
        for (int iy = 0; iy < 4; iy++)
          for (int ix = 0; ix < 8; ix++) {
            if (ix == 0 && iy == 0) continue;
            block[iy * 8 + ix] = co[(y + iy * 2) * 8 + ix];
          }
        scaled_idct(4, 8, block, px + (size_t)y * 4 * stride, stride, scratch);
      }
      break;
    }
    case 3: { /* DCT4X4 (541-568) */
      float b00 = co[0], b01 = co[1], b10 = co[8], b11 = co[9];
      float dcs[4];
      dcs[0] = b00 + b01 + b10 + b11;
      dcs[1] = b00 + b01 - b10 - b11;
      dcs[2] = b00 - b01 + b10 - b11;
      dcs[3] = b00 - b01 - b10 + b11;
      for (int y = 0; y < 2; y++)
        for (int x = 0; x < 2; x++) {
          float block[16];
          block[0] = dcs[y * 2 + x];
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 0 && iy == 0) continue;
              block[iy * 4 + ix] = co[(y + iy * 2) * 8 + x + ix * 2];
            }
          scaled_idct(4, 4, block, px + (size_t)y * 4 * stride + x * 4, stride, scratch);
        }
      break;
    }
    case 2: { /* DCT2X2 (569-581) */
      float c[64];
      memcpy(c, co, sizeof(c));
      idct2_top_block(2, c, c);
      idct2_top_block(4, c, c);
      idct2_top_block(8, c, c);
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) px[(size_t)y * stride + x] = c[y * 8 + x];
      break;
    }
    case 14: case 15: case 16: case 17:
      afv_to_pixels(strategy - 14, co, px, stride);
      break;
  }
}

/* LowestFrequenciesFromDC / ReinterpretingDCT (dec_transforms-inl.h:35-64,691-818).
 * Output row stride = max(cy,cx)*8; for 8x8-class strategies llf[0] = dc[0]. */
static void llf_from_dc(int strategy, const float* dc, size_t dc_stride, float* llf, float* scratch) {
  const int cy = kRows8[strategy], cx = kCols8[strategy];
  if (!is_plain_dct(strategy) || (cy == 1 && cx == 1)) {
    llf[0] = dc[0];
    return;
  }
  float* block = scratch;                    /* cy*cx */
  float* s2 = scratch + (size_t)cy * cx;     /* cy*cx + 2*32 */
  scaled_dct(cy, cx, dc, dc_stride, block, s2);
  const size_t os = (size_t)(cy > cx ? cy : cx) * 8;
  const float* sy = RESAMPLE(cy);
  const float* sx = RESAMPLE(cx);
  if (cy < cx) {
    for (int y = 0; y < cy; y++)
      for (int x = 0; x < cx; x++) llf[y * os + x] = block[y * cx + x] * sy[y] * sx[x];
  } else {
    for (int y = 0; y < cx; y++)
      for (int x = 0; x < cy; x++) llf[y * os + x] = block[y * cy + x] * sx[y] * sy[x];
  }
}

/* ------------------------------------------------------------------------- */
/* public function-level entry points                                         */
/* ------------------------------------------------------------------------- */
int jxo_transform_to_pixels(int strategy, const float* coeffs, float* pixels, size_t stride) {
  if (strategy < 0 || strategy >= 27) return 1;
  size_t n = (size_t)64 * kRows8[strategy] * kCols8[strategy];
  float* buf = (float*)malloc((3 * n + 1024) * sizeof(float));
  if (!buf) return 2;
  memcpy(buf, coeffs, n * sizeof(float));
  transform_to_pixels(strategy, buf, pixels, stride, buf + n);
  free(buf);
  return 0;
}

int jxo_llf_from_dc(int strategy, const float* dc, size_t dc_stride, float* block) {
  if (strategy < 0 || strategy >= 27) return 1;
  float scratch[2 * 1024 + 128];
  llf_from_dc(strategy, dc, dc_stride, block, scratch);
  return 0;
}

int jxo_scaled_dct(int rows, int cols, const float* px, size_t stride, float* out) {
  float* scratch = (float*)malloc(((size_t)rows * cols + 1024) * sizeof(float));
  if (!scratch) return 2;
  scaled_dct(rows, cols, px, stride, out, scratch);
  free(scratch);
  return 0;
}

/* AdjustQuantBias (lib/jxl/quantizer-inl.h:35-67).
 * rcp_mode 0: correctly rounded reciprocal (what the spec's q - b3/q means up to 1 ulp
 *             of the bias term; this is what the CUDA path computes).
 * rcp_mode 1: the host's rcpss, i.e. Highway's ApproximateReciprocal on the AVX2 target
 *             (third_party/highway/hwy/ops/x86_256-inl.h:2673-2675) -- reproduces the
 *             reference bit for bit on the same CPU family; used to pin this file. */
float jxo_adjust_quant_bias(int c, int32_t q, const float* biases, int rcp_mode) {
  const float quant = (float)q;
  const float aq = fabsf(quant);
  if (aq < 1.125f) {
    if (!(aq > 0.0f)) return 0.0f;
    return q < 0 ? -biases[c] : biases[c];
  }
  float r;
#if defined(__x86_64__)
  if (rcp_mode == 1) {
    r = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(quant)));
  } else
#endif
  {
    r = 1.0f / quant;
  }
  return fmaf(-biases[3], r, quant);
}

/* ------------------------------------------------------------------------- */
/* frame level                                                                */
/* ------------------------------------------------------------------------- */
static inline int64_t mirror(int64_t x, int64_t size) {
  /* Mirror (lib/jxl/image_ops.h:184-196) */
  while (x < 0 || x >= size) {
    if (x < 0) x = -x - 1;
    else x = 2 * size - 1 - x;
  }
  return x;
}

uint32_t jxo_effective_stage_mask(const jxlgpu_frame* f) {
  if (f->stage_mask & JXLGPU_STAGE_EXPLICIT) return f->stage_mask & 63u;
  uint32_t m = JXLGPU_STAGE_XYB | (f->stage_mask & JXLGPU_STAGE_SRGB); /* order: dec_cache.cc:151-170,259-260 */
  if (f->gab) m |= JXLGPU_STAGE_GAB;
  if (f->epf_iters >= 3) m |= JXLGPU_STAGE_EPF0;
  if (f->epf_iters >= 1) m |= JXLGPU_STAGE_EPF1;
  if (f->epf_iters >= 2) m |= JXLGPU_STAGE_EPF2;
  return m;
}

/* ComputeSigma (lib/jxl/epf.cc:39-133) for the whole frame. sigma: (yb+4) x (xb+4)
 * floats holding 1/sigma at [by+2][bx+2], one mirrored block of border. */
void jxo_compute_sigma(const jxlgpu_frame* f, float* sigma) {
  const size_t xb = f->xsize_blocks, yb = f->ysize_blocks, ss = xb + 4;
  const float kInvSigmaNum = -1.1715728752538099024f; /* epf.h:19 */
  for (size_t i = 0; i < (yb + 4) * ss; i++) sigma[i] = 0.0f;
  for (size_t by = 0; by < yb; by++)
    for (size_t bx = 0; bx < xb; bx++) {
      uint8_t raw = f->ac_strategy[by * f->ac_strategy_stride + bx];
      if (!(raw & 1)) continue;
      int s = raw >> 1;
      int32_t quant = f->raw_quant[by * f->raw_quant_stride + bx];
      float sigma_quant = f->epf_quant_mul / (f->quant_scale * quant * kInvSigmaNum);
      for (int iy = 0; iy < kRows8[s]; iy++)
        for (int ix = 0; ix < kCols8[s]; ix++) {
          uint8_t sharp = f->epf_sharpness[(by + iy) * f->epf_sharpness_stride + bx + ix];
          float sg = sigma_quant * f->epf_sharp_lut[sharp];
          sg = sg < -1e-4f ? sg : -1e-4f; /* std::min(-1e-4f, sigma) */
          sigma[(by + iy + 2) * ss + bx + ix + 2] = 1.0f / sg;
        }
    }
  /* one mirrored block of border (epf.cc:84-130) */
  for (size_t y = 2; y < yb + 2; y++) {
    sigma[y * ss + 1] = sigma[y * ss + 2];
    sigma[y * ss + xb + 2] = sigma[y * ss + xb + 1];
  }
  for (size_t x = 1; x < xb + 3; x++) {
    sigma[1 * ss + x] = sigma[2 * ss + x];
    sigma[(yb + 2) * ss + x] = sigma[(yb + 1) * ss + x];
  }
}

/* DecodeGroupImpl draw branch + DequantBlock/DequantLane (dec_group.cc:115-181,262-457):
 * one AC group -> post-IDCT XYB pixels in planes[c] (row stride ps). */
static void dequant_idct_group(const jxlgpu_frame* f, const void* const coeff[3], size_t g,
                               int rcp_mode, float* const planes[3], size_t ps, float* work) {
  const size_t xg = (f->xsize_blocks + 31) / 32;
  const size_t gx = g % xg, gy = g / xg;
  const size_t bx0 = gx * 32, by0 = gy * 32;
  const size_t nbx = f->xsize_blocks - bx0 < 32 ? f->xsize_blocks - bx0 : 32;
  const size_t nby = f->ysize_blocks - by0 < 32 ? f->ysize_blocks - by0 : 32;
  float* block = work;                 /* 3 * 65536 */
  float* scratch = work + 3 * 65536;   /* 2 * 65536 + 1024 */
  size_t offset = 0;
  for (size_t by = 0; by < nby; by++) {
    const size_t ty = (by0 + by) / 8;
    for (size_t bx = 0; bx < nbx; bx++) {
      const size_t abx = bx0 + bx, aby = by0 + by;
      uint8_t raw = f->ac_strategy[aby * f->ac_strategy_stride + abx];
      if (!(raw & 1)) continue;
      const int kind = raw >> 1;
      const size_t tx = abx / 8;
      const size_t size = (size_t)64 * kRows8[kind] * kCols8[kind];
      const int32_t quant = f->raw_quant[aby * f->raw_quant_stride + abx];
      /* ColorCorrelation::YtoXRatio/YtoBRatio (chroma_from_luma.h:51-57) */
      const float x_cc = f->cfl_base_x + f->ytox_map[ty * f->cmap_stride + tx] * f->cfl_color_scale;
      const float b_cc = f->cfl_base_b + f->ytob_map[ty * f->cmap_stride + tx] * f->cfl_color_scale;
      const float s = f->inv_global_scale / quant;
      const float sx = s * f->x_dm_multiplier, sy = s, sb = s * f->b_dm_multiplier;
      const float* mx = f->dequant_table + f->dequant_offsets[3 * kind + 0];
      const float* my = f->dequant_table + f->dequant_offsets[3 * kind + 1];
      const float* mb = f->dequant_table + f->dequant_offsets[3 * kind + 2];
      for (size_t k = 0; k < size; k++) {
        int32_t qx, qy, qb;
        if (f->ac_type == JXLGPU_AC_INT16) {
          qx = ((const int16_t*)coeff[0])[g * 65536 + offset + k];
          qy = ((const int16_t*)coeff[1])[g * 65536 + offset + k];
          qb = ((const int16_t*)coeff[2])[g * 65536 + offset + k];
        } else {
          qx = ((const int32_t*)coeff[0])[g * 65536 + offset + k];
          qy = ((const int32_t*)coeff[1])[g * 65536 + offset + k];
          qb = ((const int32_t*)coeff[2])[g * 65536 + offset + k];
        }
        const float x_mul = mx[k] * sx, y_mul = my[k] * sy, b_mul = mb[k] * sb;
        const float dx = jxo_adjust_quant_bias(0, qx, f->quant_biases, rcp_mode) * x_mul;
        const float dy = jxo_adjust_quant_bias(1, qy, f->quant_biases, rcp_mode) * y_mul;
        const float db = jxo_adjust_quant_bias(2, qb, f->quant_biases, rcp_mode) * b_mul;
        block[k] = fmaf(x_cc, dy, dx);
        block[size + k] = dy;
        block[2 * size + k] = fmaf(b_cc, dy, db);
      }
      offset += size;
      for (int c = 0; c < 3; c++)
        llf_from_dc(kind, f->dc[c] + aby * f->dc_stride + abx, f->dc_stride, block + c * size,
                    scratch);
      for (int c = 0; c < 3; c++)
        transform_to_pixels(kind, block + c * size, planes[c] + aby * 8 * ps + abx * 8, ps, scratch);
    }
  }
}

/* GaborishStage (render_pipeline/stage_gaborish.cc:31-100) over the full frame with
 * mirrored borders about the true image size (low_memory_render_pipeline.cc:475-517,
 * simple_render_pipeline.cc:129-164). */
static void gaborish(const jxlgpu_frame* f, float* const in[3], float* const out[3], size_t ps) {
  const int64_t W = f->xsize, H = f->ysize;
  for (int c = 0; c < 3; c++) {
    float w0 = 1.0f, w1 = f->gab_weights[2 * c], w2 = f->gab_weights[2 * c + 1];
    const float div = w0 + 4 * (w1 + w2);
    const float mul = 1.0f / div;
    w0 *= mul; w1 *= mul; w2 *= mul;
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < H; y++) {
      const float* rt = in[c] + mirror(y - 1, H) * ps;
      const float* rm = in[c] + y * ps;
      const float* rb = in[c] + mirror(y + 1, H) * ps;
      float* ro = out[c] + y * ps;
      for (int64_t x = 0; x < W; x++) {
        const int64_t xl = mirror(x - 1, W), xr = mirror(x + 1, W);
        const float sum1 = (rm[xl] + rm[xr]) + (rt[x] + rb[x]);
        const float sum2 = (rt[xl] + rt[xr]) + (rb[xl] + rb[xr]);
        ro[x] = fmaf(sum2, w2, fmaf(sum1, w1, rm[x] * w0));
      }
    }
  }
}

typedef struct {
  const float* p[3];
  size_t ps;
  int64_t W, H;
} planes_t;

static inline float px(const planes_t* im, int c, int64_t y, int64_t x) {
  return im->p[c][mirror(y, im->H) * im->ps + mirror(x, im->W)];
}

static const float kMinSigma = -3.90524291751269967465540850526868f; /* epf.h:22 */

static inline float epf_weight(float sad, float inv_sigma) {
  /* Weight (stage_epf.cc:47-50): ZeroIfNegative(MulAdd(sad, inv_sigma, 1)) */
  float v = fmaf(sad, inv_sigma, 1.0f);
  return v < 0.0f ? 0.0f : v;
}

/* which = 0/1/2: EPF0Stage / EPF1Stage / EPF2Stage (stage_epf.cc:54-193,197-379,383-506) */
static void epf(const jxlgpu_frame* f, int which, const float* sigma, float* const in[3],
                float* const out[3], size_t ps) {
  planes_t im = {{in[0], in[1], in[2]}, ps, f->xsize, f->ysize};
  const size_t ss = f->xsize_blocks + 4;
  float sm;
  if (which == 0) sm = (float)(f->epf_pass0_sigma_scale * 1.65);
  else if (which == 1) sm = 1.65f;
  else sm = (float)(f->epf_pass2_sigma_scale * 1.65);
  const float bsm = sm * f->epf_border_sad_mul;
  static const int sads_off[12][2] = {{-2, 0}, {-1, -1}, {-1, 0}, {-1, 1}, {0, -2}, {0, -1},
                                      {0, 1},  {0, 2},   {1, -1}, {1, 0},  {1, 1},  {2, 0}};
  static const int plus_off[5][2] = {{0, 0}, {-1, 0}, {0, -1}, {1, 0}, {0, 1}};
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < im.H; y++) {
    const float* row_sigma = sigma + (y / 8 + 2) * ss;
    const int border_row = (y % 8 == 0 || y % 8 == 7);
    for (int64_t x = 0; x < im.W; x++) {
      const float s = row_sigma[x / 8 + 2];
      if (s < kMinSigma) {
        for (int c = 0; c < 3; c++) out[c][y * ps + x] = in[c][y * ps + x];
        continue;
      }
      const int ix = (int)(x % 8);
      const float vsm = (border_row || ix == 0 || ix == 7) ? bsm : sm;
      const float inv_sigma = s * vsm;
      float w = 1.0f;
      float X = in[0][y * ps + x], Y = in[1][y * ps + x], B = in[2][y * ps + x];
      if (which == 0) {
        float sads[12];
        for (int i = 0; i < 12; i++) sads[i] = 0.0f;
        for (int c = 0; c < 3; c++) {
          const float scale = f->epf_channel_scale[c];
          for (int i = 0; i < 12; i++) {
            float sad = 0.0f;
            for (int o = 0; o < 5; o++) {
              const float r11 = px(&im, c, y + plus_off[o][0], x + plus_off[o][1]);
              const float c11 = px(&im, c, y + sads_off[i][0] + plus_off[o][0],
                                   x + sads_off[i][1] + plus_off[o][1]);
              sad = sad + fabsf(r11 - c11);
            }
            sads[i] = fmaf(sad, scale, sads[i]);
          }
        }
        for (int i = 0; i < 12; i++) {
          const float wt = epf_weight(sads[i], inv_sigma);
          const int64_t yy = y + sads_off[i][0], xx = x + sads_off[i][1];
          w = w + wt;
          X = fmaf(wt, px(&im, 0, yy, xx), X);
          Y = fmaf(wt, px(&im, 1, yy, xx), Y);
          B = fmaf(wt, px(&im, 2, yy, xx), B);
        }
      } else if (which == 1) {
        float sad0 = 0, sad1 = 0, sad2 = 0, sad3 = 0;
        for (int c = 0; c < 3; c++) {
          /* pXY: X = column 0..4, Y = row 0..4, centre p22 (stage_epf.cc:278-336) */
#define P(col, row) px(&im, c, y + (row) - 2, x + (col) - 2)
          const float p20 = P(2, 0), p21 = P(2, 1), p11 = P(1, 1), p31 = P(3, 1);
          const float p02 = P(0, 2), p12 = P(1, 2), p22 = P(2, 2), p32 = P(3, 2), p42 = P(4, 2);
          const float p13 = P(1, 3), p23 = P(2, 3), p33 = P(3, 3), p24 = P(2, 4);
#undef P
          float t;
          float sad0c = fabsf(p20 - p21);
          float sad1c = fabsf(p11 - p21);
          float sad2c = fabsf(p31 - p21);
          sad1c = sad1c + fabsf(p02 - p12);
          sad0c = sad0c + fabsf(p11 - p12);
          t = fabsf(p12 - p22);
          sad1c = sad1c + t;
          sad2c = sad2c + t;
          t = fabsf(p22 - p21);
          float sad3c = t;
          sad0c = sad0c + t;
          sad0c = sad0c + fabsf(p31 - p32);
          t = fabsf(p22 - p32);
          sad1c = sad1c + t;
          sad2c = sad2c + t;
          sad2c = sad2c + fabsf(p42 - p32);
          sad3c = sad3c + fabsf(p13 - p12);
          t = fabsf(p22 - p23);
          sad0c = sad0c + t;
          sad3c = sad3c + t;
          sad1c = sad1c + fabsf(p13 - p23);
          sad2c = sad2c + fabsf(p33 - p23);
          sad3c = sad3c + fabsf(p33 - p32);
          sad3c = sad3c + fabsf(p24 - p23);
          const float scale = f->epf_channel_scale[c];
          sad0 = fmaf(sad0c, scale, sad0);
          sad1 = fmaf(sad1c, scale, sad1);
          sad2 = fmaf(sad2c, scale, sad2);
          sad3 = fmaf(sad3c, scale, sad3);
        }
        const float sd[4] = {sad0, sad1, sad2, sad3};
        static const int off[4][2] = {{-1, 0}, {0, -1}, {0, 1}, {1, 0}};
        for (int i = 0; i < 4; i++) {
          const float wt = epf_weight(sd[i], inv_sigma);
          const int64_t yy = y + off[i][0], xx = x + off[i][1];
          w = w + wt;
          X = fmaf(wt, px(&im, 0, yy, xx), X);
          Y = fmaf(wt, px(&im, 1, yy, xx), Y);
          B = fmaf(wt, px(&im, 2, yy, xx), B);
        }
      } else {
        const float rx = X, ry = Y, rb = B;
        static const int off[4][2] = {{-1, 0}, {0, -1}, {0, 1}, {1, 0}};
        for (int i = 0; i < 4; i++) {
          const int64_t yy = y + off[i][0], xx = x + off[i][1];
          const float cx = px(&im, 0, yy, xx), cy = px(&im, 1, yy, xx), cb = px(&im, 2, yy, xx);
          float sad = fabsf(cx - rx) * f->epf_channel_scale[0];
          sad = fmaf(fabsf(cy - ry), f->epf_channel_scale[1], sad);
          sad = fmaf(fabsf(cb - rb), f->epf_channel_scale[2], sad);
          const float wt = epf_weight(sad, inv_sigma);
          w = w + wt;
          X = fmaf(wt, cx, X);
          Y = fmaf(wt, cy, Y);
          B = fmaf(wt, cb, B);
        }
      }
      const float inv_w = 1.0f / w; /* JXL_HIGH_PRECISION: Div(1, w) (stage_epf.cc:170-171) */
      out[0][y * ps + x] = X * inv_w;
      out[1][y * ps + x] = Y * inv_w;
      out[2][y * ps + x] = B * inv_w;
    }
  }
}

/* XybToRgb (lib/jxl/dec_xyb-inl.h:38-86), in place */
/* kYCbCrStage (lib/jxl/render_pipeline/stage_ycbcr.cc:33-71): full-range BT.601, planes 0 = Cb, 1 = Y, 2 = Cr */
static void ycbcr_to_rgb(const jxlgpu_frame* f, float* const p[3], size_t ps) {
  const float c128 = 128.0f / 255, crcr = 1.402f, cgcb = -0.114f * 1.772f / 0.587f, cgcr = -0.299f * 1.402f / 0.587f,
              cbcb = 1.772f;
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)f->ysize; y++)
    for (size_t x = 0; x < f->xsize; x++) {
      const size_t i = (size_t)y * ps + x;
      const float yv = p[1][i] + c128, cb = p[0][i], cr = p[2][i];
      p[0][i] = fmaf(crcr, cr, yv);
      p[1][i] = fmaf(cgcr, cr, fmaf(cgcb, cb, yv));
      p[2][i] = fmaf(cbcb, cb, yv);
    }
}

static void xyb_to_linear(const jxlgpu_frame* f, float* const p[3], size_t ps) {
  if (f->color_transform == 1) { ycbcr_to_rgb(f, p, ps); return; }
  const float* m = f->inverse_opsin_matrix;
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)f->ysize; y++)
    for (size_t x = 0; x < f->xsize; x++) {
      const size_t i = (size_t)y * ps + x;
      const float ox = p[0][i], oy = p[1][i], ob = p[2][i];
      float gr = oy + ox, gg = oy - ox, gb = ob;
      gr = gr - f->opsin_biases_cbrt[0];
      gg = gg - f->opsin_biases_cbrt[1];
      gb = gb - f->opsin_biases_cbrt[2];
      const float r2 = gr * gr, g2 = gg * gg, b2 = gb * gb;
      const float mr = fmaf(r2, gr, f->opsin_biases[0]);
      const float mg = fmaf(g2, gg, f->opsin_biases[1]);
      const float mb = fmaf(b2, gb, f->opsin_biases[2]);
      float lr = m[0] * mr, lg = m[3] * mr, lb = m[6] * mr;
      lr = fmaf(m[1], mg, lr); lg = fmaf(m[4], mg, lg); lb = fmaf(m[7], mg, lb);
      lr = fmaf(m[2], mb, lr); lg = fmaf(m[5], mb, lg); lb = fmaf(m[8], mb, lb);
      p[0][i] = lr; p[1][i] = lg; p[2][i] = lb;
    }
}

/* ---- the DC stage in front of the hot path (not yet on the GPU) ---- */

/* DequantDC, 4:4:4 branch (lib/jxl/compressed_dc.cc:199-232): q[c] = quantised DC planes (X, Y, B),
 * fac_c = dc_factors[c] * mul (one float product, as `Set(df, dc_factors[0] * mul)`), Y stored as is,
 * X and B get the chroma-from-luma term with one FMA each.  out: [3][ys][xs]. */
void jxo_dequant_dc(const int32_t* const q[3], size_t xs, size_t ys, const float* dc_factors, float mul,
                    const float* cfl_factors, float* out) {
  const float fac_x = dc_factors[0] * mul, fac_y = dc_factors[1] * mul, fac_b = dc_factors[2] * mul;
  const size_t n = xs * ys;
  for (size_t i = 0; i < n; i++) {
    const float in_x = (float)q[0][i] * fac_x;
    const float in_y = (float)q[1][i] * fac_y;
    const float in_b = (float)q[2][i] * fac_b;
    out[n + i] = in_y;
    out[i] = fmaf(in_y, cfl_factors[0], in_x);
    out[2 * n + i] = fmaf(in_y, cfl_factors[2], in_b);
  }
}

/* AdaptiveDCSmoothing (lib/jxl/compressed_dc.cc:50-197): 3x3 smoothing of the interior, weighted by how
 * far (in quantisation steps) the smoothed value is from the original in the worst channel.
 * dc: [3][ys][xs] in place; images with a side <= 2 are left alone (:133). */
int jxo_adaptive_dc_smoothing(const float* dc_factors, float* dc, size_t xs, size_t ys) {
  if (ys <= 2 || xs <= 2) return 0;
  const float w1 = 0.20345139757231578f, w2 = 0.0334829185968739f;
  const float w0 = 1.0f - 4.0f * (w1 + w2);
  const size_t n = xs * ys;
  float* sm_out = (float*)malloc(3 * n * sizeof(float));
  if (!sm_out) return 2;
  memcpy(sm_out, dc, 3 * n * sizeof(float)); /* borders stay (:143-148, :170-174) */
#pragma omp parallel for schedule(static)
  for (int64_t y = 1; y < (int64_t)ys - 1; y++)
    for (size_t x = 1; x + 1 < xs; x++) {
      float mc[3], sm[3];
      float gap = 0.5f;
      for (int c = 0; c < 3; c++) {
        const float* p = dc + (size_t)c * n + (size_t)y * xs + x;
        const float tl = p[-(ptrdiff_t)xs - 1], tc = p[-(ptrdiff_t)xs], tr = p[-(ptrdiff_t)xs + 1];
        const float ml = p[-1], mr = p[1];
        const float bl = p[xs - 1], bc = p[xs], br = p[xs + 1];
        mc[c] = p[0];
        const float corner = (tl + tr) + (bl + br);
        const float side = (ml + mr) + (tc + bc);
        sm[c] = fmaf(corner, w2, fmaf(side, w1, mc[c] * w0));
        const float g = fabsf((mc[c] - sm[c]) / dc_factors[c]);
        gap = gap > g ? gap : g; /* Max(gap, g): maxps(a, b) = a > b ? a : b */
      }
      float factor = fmaf(-4.0f, gap, 3.0f);
      if (factor < 0.0f) factor = 0.0f; /* ZeroIfNegative */
      for (int c = 0; c < 3; c++) sm_out[(size_t)c * n + (size_t)y * xs + x] = fmaf(sm[c] - mc[c], factor, mc[c]);
    }
  memcpy(dc, sm_out, 3 * n * sizeof(float));
  free(sm_out);
  return 0;
}

/* TF_SRGB::EncodedFromDisplay (lib/jxl/cms/transfer_functions-inl.h:244-267): the branch OpRgb takes
 * with JXL_HIGH_PRECISION (stage_from_linear.cc:42-53, common.h:15-16).  Rational polynomial in
 * sqrt(|x|), Horner with fused multiply-adds (rational_polynomial-inl.h:59-97), true division
 * (FastDivision<float>, :36-52), linear segment below 0.0031308, sign carried over. */
float jxo_srgb_from_linear(float v) {
  static const float p[5] = {-5.135152395e-04f, 5.287254571e-03f, 3.903842876e-01f, 1.474205315e+00f,
                             7.352629620e-01f};
  static const float q[5] = {1.004519624e-02f, 3.036675394e-01f, 1.340816930e+00f, 9.258482155e-01f,
                             2.424867759e-02f};
  const float x = fabsf(v);
  const float linear = x * 12.92f;
  const float s = sqrtf(x);
  float yp = p[4], yq = q[4];
  for (int i = 3; i >= 0; i--) {
    yp = fmaf(yp, s, p[i]);
    yq = fmaf(yq, s, q[i]);
  }
  const float poly = yp / yq;
  const float mag = x > 0.0031308f ? poly : linear;
  return copysignf(fabsf(mag), v);
}

/* MakeUnsigned (lib/jxl/render_pipeline/stage_write.cc:455-479): scale, ordered dither for 8-bit
 * (x/y/channel offsets :466-471; the 48-wide padded rows make the lane index wrap mod 32), clamp
 * (Min(Max(v, 0), mul); a NaN becomes 0 as with maxps), round half to even (NearestInt). */
uint32_t jxo_make_unsigned(float v, int bits, size_t x, size_t y, int c) {
  const float mul = (float)((1u << bits) - 1u);
  v = v * mul;
  if (bits == 8) v = v + JXT_DITHER[(y + 13 * (size_t)c) & 31][(x + 23 * (size_t)c) & 31];
  float t = v > 0.0f ? v : 0.0f;
  t = t < mul ? t : mul;
  return (uint32_t)(int32_t)nearbyintf(t);
}

/* DemoteTo(float16) as StoreFloat16Row uses it (stage_write.cc:590-640): IEEE round-to-nearest-even
 * (vcvtps2ph with rounding control 0), overflow to infinity, gradual underflow, NaN stays NaN. */
uint16_t jxo_f16_from_f32(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t a = u & 0x7fffffffu;
  if (a >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (a > 0x7f800000u ? 0x200u | ((a >> 13) & 0x3ffu) : 0u));
  if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); /* >= 65520 rounds to infinity */
  if (a < 0x33000001u) return (uint16_t)sign;               /* <= 2^-25 rounds to zero (tie to even) */
  const int e = (int)(a >> 23) - 127;
  uint32_t m = (a & 0x7fffffu) | 0x800000u;
  int shift = 13;
  uint32_t h_exp = (uint32_t)(e + 15);
  if (e < -14) { /* subnormal half */
    shift += -14 - e;
    h_exp = 0;
  }
  const uint32_t halfway = 1u << (shift - 1);
  const uint32_t rem = m & ((1u << shift) - 1u);
  uint32_t r = m >> shift;
  if (rem > halfway || (rem == halfway && (r & 1u))) r++;
  /* r carries the implicit bit for normals: adding it to (h_exp - 1) << 10 lets a mantissa carry
   * roll into the exponent */
  const uint32_t h = h_exp ? ((h_exp - 1u) << 10) + r : r;
  return (uint16_t)(sign | h);
}

/* FromLinearStage<OpRgb> over the whole frame, in place (stage_from_linear.cc:86-96) */
static void srgb_from_linear(const jxlgpu_frame* f, float* const p[3], size_t ps) {
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)f->ysize; y++)
    for (int c = 0; c < 3; c++)
      for (size_t x = 0; x < f->xsize; x++) p[c][(size_t)y * ps + x] = jxo_srgb_from_linear(p[c][(size_t)y * ps + x]);
}

size_t jxo_out_bytes(const jxlgpu_frame* f) {
  const size_t n = (size_t)f->xsize * f->ysize;
  switch (f->out_format) {
    case JXLGPU_OUT_RGB_U8: return n * 3;
    case JXLGPU_OUT_RGBA_U8: return n * 4;
    case JXLGPU_OUT_RGB_U16: case JXLGPU_OUT_RGB_F16: return n * 6;
    default: return n * 12;
  }
}

/* ---- UpsamplingStage (lib/jxl/render_pipeline/stage_upsampling.cc:51-271), SURVEY.md §8f rank 4 ----
 * N = 2, 4, 8.  kernel[k][i], k = N*oy + ox, i = 5*(iy+2) + (ix+2): the 25 taps of output sub-pixel (ox, oy),
 * expanded from the N/2 x N/2 x 25 symmetric weights exactly as the stage's constructor does (:61-86).
 * Per input pixel and sub-pixel (:246-262): three accumulators over taps i, i+1, i+2 (Mul for the first three
 * taps, MulAdd afterwards), acc0 takes tap 24, result = (acc1 + acc2) + acc0, clamped to the minimum / maximum of
 * the 5x5 input window (:152-206).  Input mirrored about the coded frame size. */
void jxo_upsampling_kernel(int N, const float* weights, float* kernel /* N*N*25 */) {
  const int H = N / 2;
  for (int ky = 0; ky < H; ky++)
    for (int kx = 0; kx < H; kx++) {
      const int o0 = (ky * N + kx) * 25, o1 = (ky * N + (N - 1 - kx)) * 25;
      const int o2 = ((N - 1 - ky) * N + kx) * 25, o3 = ((N - 1 - ky) * N + (N - 1 - kx)) * 25;
      for (int py = 0; py < 5; py++)
        for (int px = 0; px < 5; px++) {
          const int j = 5 * ky + py, i = 5 * kx + px;
          const int my = i < j ? i : j, mx = i < j ? j : i;
          const float w = weights[5 * H * my - my * (my - 1) / 2 + mx - my];
          kernel[o0 + py * 5 + px] = w;
          kernel[o1 + py * 5 + (4 - px)] = w;
          kernel[o2 + (4 - py) * 5 + px] = w;
          kernel[o3 + (4 - py) * 5 + (4 - px)] = w;
        }
    }
}

static size_t mirror_sz(int64_t x, int64_t size) {
  while (x < 0 || x >= size) x = x < 0 ? -x - 1 : 2 * size - 1 - x;
  return (size_t)x;
}

/* in: w x h (row stride ps_in); out: ow x oh (ow <= N*w, oh <= N*h; row stride ps_out) */
void jxo_upsample_plane(int N, const float* kernel, const float* in, size_t w, size_t h, size_t ps_in, float* out,
                        size_t ow, size_t oh, size_t ps_out) {
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)h; y++)
    for (size_t x = 0; x < w; x++) {
      float v[25];
      float mn = 0, mx = 0;
      for (int iy = -2; iy <= 2; iy++)
        for (int ix = -2; ix <= 2; ix++) {
          const float p = in[mirror_sz(y + iy, (int64_t)h) * ps_in + mirror_sz((int64_t)x + ix, (int64_t)w)];
          v[5 * (iy + 2) + ix + 2] = p;
          if (iy == -2 && ix == -2) { mn = mx = p; }
          else { mn = p < mn ? p : mn; mx = p > mx ? p : mx; }
        }
      for (int oy = 0; oy < N; oy++)
        for (int ox = 0; ox < N; ox++) {
          const size_t X = x * (size_t)N + ox, Y = (size_t)y * N + oy;
          if (X >= ow || Y >= oh) continue;
          const float* k = kernel + (N * oy + ox) * 25;
          float a0 = v[0] * k[0], a1 = v[1] * k[1], a2 = v[2] * k[2];
          for (int i = 3; i < 24; i += 3) {
            a0 = fmaf(v[i], k[i], a0);
            a1 = fmaf(v[i + 1], k[i + 1], a1);
            a2 = fmaf(v[i + 2], k[i + 2], a2);
          }
          a0 = fmaf(v[24], k[24], a0);
          float r = (a1 + a2) + a0;
          r = r < mn ? mn : r;  /* Clamp(v, lo, hi) = Min(Max(lo, v), hi) */
          r = r > mx ? mx : r;
          out[Y * ps_out + X] = r;
        }
    }
}

/* ---- noise: Random3Planes + ConvolveNoiseStage + AddNoiseStage ----
 * lib/jxl/dec_noise.cc:45-110,120-152 (generation per 256x256 tile of the OUTPUT image, Xorshift128Plus seeded with
 * (visible_frame_index, nonvisible_frame_index, x0, y0), lib/jxl/xorshift128plus-inl.h:31-91),
 * lib/jxl/render_pipeline/stage_noise.cc:263-304 (convolution) and :140-251 (strength LUT, mixing into X, Y, B). */
typedef struct { uint64_t s0[8], s1[8]; } jxo_rng;
static uint64_t splitmix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void rng_init(jxo_rng* r, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  r->s0[0] = splitmix64((((uint64_t)a << 32) + b) + 0x9E3779B97F4A7C15ull);
  r->s1[0] = splitmix64((((uint64_t)c << 32) + d) + 0x9E3779B97F4A7C15ull);
  for (int i = 1; i < 8; i++) {
    r->s0[i] = splitmix64(r->s0[i - 1]);
    r->s1[i] = splitmix64(r->s1[i - 1]);
  }
}
static void rng_fill(jxo_rng* r, uint64_t* bits) {
  for (int i = 0; i < 8; i++) {
    uint64_t s1 = r->s0[i];
    const uint64_t s0 = r->s1[i];
    bits[i] = s1 + s0;
    r->s0[i] = s0;
    s1 ^= s1 << 23;
    s1 ^= s0 ^ (s1 >> 18) ^ (s0 >> 5);
    r->s1[i] = s1;
  }
}
/* three planes W x H of floats in [1, 2) (row stride W) */
void jxo_noise_planes(uint32_t visible, uint32_t nonvisible, size_t W, size_t H, float* planes) {
  for (size_t y0 = 0; y0 < H; y0 += 256)
    for (size_t x0 = 0; x0 < W; x0 += 256) {
      const size_t xs = W - x0 < 256 ? W - x0 : 256, ys = H - y0 < 256 ? H - y0 : 256;
      jxo_rng rng;
      rng_init(&rng, visible, nonvisible, (uint32_t)x0, (uint32_t)y0);
      for (int p = 0; p < 3; p++)
        for (size_t y = 0; y < ys; y++) {
          float* row = planes + (size_t)p * W * H + (y0 + y) * W + x0;
          uint64_t b64[8];
          uint32_t b32[16];
          size_t x = 0;
          for (; x + 16 < xs; x += 16) {  /* only entire batches */
            rng_fill(&rng, b64);
            memcpy(b32, b64, sizeof(b32));
            for (int i = 0; i < 16; i++) {
              const uint32_t u = (b32[i] >> 9) | 0x3F800000u;
              memcpy(&row[x + i], &u, 4);
            }
          }
          rng_fill(&rng, b64);             /* the remaining pixels (<= 16) */
          memcpy(b32, b64, sizeof(b32));
          for (int i = 0; x < xs; x++, i++) {
            const uint32_t u = (b32[i] >> 9) | 0x3F800000u;
            memcpy(&row[x], &u, 4);
          }
        }
    }
}
/* ConvolveNoiseStage on one plane (mirrored borders), out-of-place */
static void noise_convolve(const float* in, float* out, size_t W, size_t H) {
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)H; y++) {
    const float* rows[5];
    for (int i = 0; i < 5; i++) rows[i] = in + mirror_sz(y + i - 2, (int64_t)H) * W;
    for (size_t x = 0; x < W; x++) {
      size_t xi[5];
      for (int i = 0; i < 5; i++) xi[i] = mirror_sz((int64_t)x + i - 2, (int64_t)W);
      float others = 0.0f;
      for (int i = 0; i < 5; i++) {
        others = others + rows[0][xi[i]];
        others = others + rows[1][xi[i]];
        others = others + rows[3][xi[i]];
        others = others + rows[4][xi[i]];
      }
      others = others + rows[2][xi[0]];
      others = others + rows[2][xi[1]];
      others = others + rows[2][xi[3]];
      others = others + rows[2][xi[4]];
      out[(size_t)y * W + x] = fmaf(others, 0.16f, rows[2][xi[2]] * -3.84f);
    }
  }
}
static float noise_strength(const float* lut, float x) {
  float scaled = x * 6.0f;   /* kNumNoisePoints - 2 */
  scaled = scaled > 0.0f ? scaled : 0.0f;
  float fl = floorf(scaled);
  float frac = scaled - fl;
  if (scaled >= 7.0f) { fl = 6.0f; frac = 1.0f; }
  const int i = (int)fl;
  float v = fmaf(lut[i + 1] - lut[i], frac, lut[i]);
  v = v < 1.0f ? v : 1.0f;   /* Clamp0ToMax: Min(x, 1) then ZeroIfNegative */
  return v < 0.0f ? 0.0f : v;
}
/* p[3]: X, Y, B planes W x H (row stride ps), in place */
void jxo_add_noise(const float* lut, float ytox, float ytob, uint32_t visible, uint32_t nonvisible, float* const p[3],
                   size_t W, size_t H, size_t ps) {
  float* raw = (float*)malloc(3 * W * H * sizeof(float));
  float* conv = (float*)malloc(3 * W * H * sizeof(float));
  if (!raw || !conv) { free(raw); free(conv); return; }
  jxo_noise_planes(visible, nonvisible, W, H, raw);
  for (int c = 0; c < 3; c++) noise_convolve(raw + (size_t)c * W * H, conv + (size_t)c * W * H, W, H);
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < (int64_t)H; y++)
    for (size_t x = 0; x < W; x++) {
      const size_t i = (size_t)y * ps + x, n = (size_t)y * W + x;
      const float vx = p[0][i], vy = p[1][i];
      const float in_g = vy - vx, in_r = vy + vx;
      const float sg = noise_strength(lut, in_g * 0.5f), sr = noise_strength(lut, in_r * 0.5f);
      const float rr = conv[n] * 0.22f, rg = conv[W * H + n] * 0.22f, rc = conv[2 * W * H + n] * 0.22f;
      const float red = sr * fmaf(0.0078125f, rr, 0.9921875f * rc);
      const float green = sg * fmaf(0.0078125f, rg, 0.9921875f * rc);
      const float sum = red + green;
      p[0][i] = fmaf(ytox, sum, red - green) + vx;
      p[1][i] = vy + sum;
      p[2][i] = fmaf(ytob, sum, p[2][i]);
    }
  free(raw); free(conv);
}

int jxo_render_frame(const jxlgpu_frame* f_in, const void* const coeff[3], int rcp_mode, void* out_v) {
  float* out = (float*)out_v;
  /* optional DC stage (quant_dc given): DequantDC per DC group + AdaptiveDCSmoothing, then as usual */
  jxlgpu_frame f_local = *f_in;
  const jxlgpu_frame* f = &f_local;
  float* dc_own = NULL;
  if (f_in->quant_dc[0]) {
    const size_t xbb = f_in->xsize_blocks, ybb = f_in->ysize_blocks, nb = xbb * ybb;
    dc_own = (float*)malloc(3 * nb * sizeof(float));
    if (!dc_own) return 2;
    const size_t xdg = (xbb + 255) / 256;
    for (size_t y = 0; y < ybb; y++)
      for (size_t x = 0; x < xbb; x++) { /* one block at a time: the group factor may change every 256 blocks */
        const int32_t qx = f_in->quant_dc[0][y * f_in->quant_dc_stride + x];
        const int32_t qy = f_in->quant_dc[1][y * f_in->quant_dc_stride + x];
        const int32_t qb = f_in->quant_dc[2][y * f_in->quant_dc_stride + x];
        const int32_t* q1[3] = {&qx, &qy, &qb};
        const float mul = f_in->dc_group_mul ? f_in->dc_group_mul[(y >> 8) * xdg + (x >> 8)] : 1.0f;
        float o[3];
        jxo_dequant_dc(q1, 1, 1, f_in->dc_factors, mul, f_in->dc_cfl_factors, o);
        for (int c = 0; c < 3; c++) dc_own[(size_t)c * nb + y * xbb + x] = o[c];
      }
    if (f_in->dc_smoothing && jxo_adaptive_dc_smoothing(f_in->dc_factors, dc_own, xbb, ybb)) { free(dc_own); return 2; }
    for (int c = 0; c < 3; c++) f_local.dc[c] = dc_own + (size_t)c * nb;
    f_local.dc_stride = xbb;
  }
  const size_t xb = f->xsize_blocks, yb = f->ysize_blocks;
  size_t ps = xb * 8;
  const size_t plane = ps * yb * 8;
  const size_t xg = (xb + 31) / 32, yg = (yb + 31) / 32;
  float* a = (float*)calloc(3 * plane, sizeof(float));
  float* b = (float*)calloc(3 * plane, sizeof(float));
  float* sigma = (float*)calloc((yb + 4) * (xb + 4), sizeof(float));
  if (!a || !b || !sigma) { free(a); free(b); free(sigma); free(dc_own); return 2; }
  float* A[3] = {a, a + plane, a + 2 * plane};
  float* B[3] = {b, b + plane, b + 2 * plane};
  int failed = 0;
#pragma omp parallel
  {
    float* work = (float*)malloc((5 * 65536 + 2048) * sizeof(float));
    if (!work) {
#pragma omp atomic write
      failed = 1;
    }
#pragma omp for schedule(dynamic)
    for (int64_t g = 0; g < (int64_t)(xg * yg); g++)
      if (work) dequant_idct_group(f, coeff, (size_t)g, rcp_mode, A, ps, work);
    free(work);
  }
  if (failed) { free(a); free(b); free(sigma); free(dc_own); return 2; }
  const uint32_t mask = jxo_effective_stage_mask(f);
  float** cur = A;
  float** nxt = B;
#define SWAP() do { float** t_ = cur; cur = nxt; nxt = t_; } while (0)
  if (mask & (JXLGPU_STAGE_EPF0 | JXLGPU_STAGE_EPF1 | JXLGPU_STAGE_EPF2)) jxo_compute_sigma(f, sigma);
  if (mask & JXLGPU_STAGE_GAB) { gaborish(f, cur, nxt, ps); SWAP(); }
  if (mask & JXLGPU_STAGE_EPF0) { epf(f, 0, sigma, cur, nxt, ps); SWAP(); }
  if (mask & JXLGPU_STAGE_EPF1) { epf(f, 1, sigma, cur, nxt, ps); SWAP(); }
  if (mask & JXLGPU_STAGE_EPF2) { epf(f, 2, sigma, cur, nxt, ps); SWAP(); }
  float* up = NULL;
  float* U[3];
  size_t ps_out = ps;
  jxlgpu_frame f_up = *f;  /* the stages after the upsampling work at the upsampled size */
  const jxlgpu_frame* fo = f;
  if (f->upsampling > 1) {
    const int N = (int)f->upsampling;
    const size_t ow = f->xsize_upsampled ? f->xsize_upsampled : (size_t)N * f->xsize;
    const size_t oh = f->ysize_upsampled ? f->ysize_upsampled : (size_t)N * f->ysize;
    float kernel[64 * 25];
    jxo_upsampling_kernel(N, f->upsampling_weights, kernel);
    up = (float*)malloc(3 * ow * oh * sizeof(float));
    if (!up) { free(a); free(b); free(sigma); free(dc_own); return 2; }
    for (int c = 0; c < 3; c++) {
      U[c] = up + (size_t)c * ow * oh;
      jxo_upsample_plane(N, kernel, cur[c], f->xsize, f->ysize, ps, U[c], ow, oh, ow);
    }
    cur = U;
    ps_out = ow;
    f_up.xsize = (uint32_t)ow;
    f_up.ysize = (uint32_t)oh;
    fo = &f_up;
  }
  if (f->noise) {  /* ConvolveNoise + AddNoise at the output resolution, before the colour transform (dec_cache.cc:232-236) */
    float* N[3] = {cur[0], cur[1], cur[2]};
    jxo_add_noise(f->noise_lut, f->cfl_base_x, f->cfl_base_b, f->visible_frame_index, f->nonvisible_frame_index, N,
                  fo->xsize, fo->ysize, ps_out);
  }
  if (mask & JXLGPU_STAGE_XYB) xyb_to_linear(fo, cur, ps_out);
  if (mask & JXLGPU_STAGE_SRGB) srgb_from_linear(fo, cur, ps_out);
#undef SWAP
  const size_t W = fo->xsize, H = fo->ysize;
  ps = ps_out;
  if (f->out_format >= JXLGPU_OUT_RGB_U8 && f->out_format <= JXLGPU_OUT_RGB_F16) {
    /* WriteToOutputStage (stage_write.cc:455-640): interleave + convert; opaque alpha = all ones */
    uint8_t* o8 = (uint8_t*)out_v;
    uint16_t* o16 = (uint16_t*)out_v;
    const uint32_t fmt = f->out_format;
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)H; y++)
      for (size_t x = 0; x < W; x++)
        for (int c = 0; c < 3; c++) {
          const float v = cur[c][(size_t)y * ps + x];
          const size_t i = (size_t)y * W + x;
          if (fmt == JXLGPU_OUT_RGB_U8) o8[i * 3 + c] = (uint8_t)jxo_make_unsigned(v, 8, x, (size_t)y, c);
          else if (fmt == JXLGPU_OUT_RGBA_U8) {
            o8[i * 4 + c] = (uint8_t)jxo_make_unsigned(v, 8, x, (size_t)y, c);
            if (c == 2) o8[i * 4 + 3] = (uint8_t)jxo_make_unsigned(1.0f, 8, x, (size_t)y, 3);
          } else if (fmt == JXLGPU_OUT_RGB_U16) o16[i * 3 + c] = (uint16_t)jxo_make_unsigned(v, 16, x, (size_t)y, c);
          else o16[i * 3 + c] = jxo_f16_from_f32(v);
        }
  } else if (f->out_format == JXLGPU_OUT_PLANAR_F32) {
    for (int c = 0; c < 3; c++)
      for (size_t y = 0; y < H; y++) memcpy(out + (c * H + y) * W, cur[c] + y * ps, W * sizeof(float));
  } else {
    for (size_t y = 0; y < H; y++)
      for (size_t x = 0; x < W; x++)
        for (int c = 0; c < 3; c++) out[(y * W + x) * 3 + c] = cur[c][y * ps + x];
  }
  free(a); free(b); free(sigma); free(dc_own); free(up);
  return 0;
}
