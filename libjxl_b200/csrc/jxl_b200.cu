// jxl_b200.cu -- context management and the C ABI of include/jxl_b200.h.
// The product path: there is NO CPU fallback in this library; without a CUDA device every
// entry point fails with JXLGPU_ERR_NO_DEVICE / JXLGPU_ERR_CUDA.
#include "../../include/jxl_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "jxl_kernels.cuh"

using namespace jxlb;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct jxlgpu_ctx {
  int device = 0;
  uint32_t num_threads = 1;
  cudaStream_t stream = nullptr;
  std::vector<cudaStream_t> up_streams;
  std::vector<cudaEvent_t> up_events;
  int num_sms = 148;
  bool in_frame = false;
  bool coeff_external = false;
  FrameDev P{};
  uint32_t num_groups = 0;
  uint32_t plan_groups = 0;      // groups the plan kernel visits (band +- halo)
  uint32_t need_g0 = 0, need_g1 = 0;
  size_t elem_size = 2;
  std::vector<uint8_t> submitted;
  DevBuf acs, quant, sharp, ytox, ytob, dc, dq, coeff[3], coeff_off, sigma, list, counts, xyb, out;
  size_t out_stride_floats = 0;
  std::atomic<uint64_t> launches{0};
  bool force_generic_filter = false;  // JXLGPU_FORCE_GENERIC_FILTER=1: tile kernel for every chain
  bool profile = false;          // record CUDA events around every kernel (bench roofline)
  cudaEvent_t prof_ev[6] = {};
  std::string last_error;
  std::mutex mu;
};

namespace {

int fail_cuda(jxlgpu_ctx* ctx, cudaError_t e, const char* what) {
  ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return e == cudaErrorMemoryAllocation ? JXLGPU_ERR_OUT_OF_MEMORY : JXLGPU_ERR_CUDA;
}

#define CU(call)                                           \
  do {                                                     \
    cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call); \
  } while (0)

size_t out_floats_per_row(const jxlgpu_frame& f) {
  return f.out_format == JXLGPU_OUT_RGB_F32 ? (size_t)f.xsize * 3 : (size_t)f.xsize;
}

// copies a strided host plane into a dense device plane
template <typename T>
cudaError_t upload_plane(void* dst, const T* src, size_t stride, size_t w, size_t h, cudaStream_t s) {
  return cudaMemcpy2DAsync(dst, w * sizeof(T), src, stride * sizeof(T), w * sizeof(T), h,
                           cudaMemcpyHostToDevice, s);
}

uint32_t effective_mask(const jxlgpu_frame& f) {
  if (f.stage_mask & JXLGPU_STAGE_EXPLICIT) return f.stage_mask & 31u;
  uint32_t m = JXLGPU_STAGE_XYB;  // PassesDecoderState::PreparePipeline order, dec_cache.cc:151-170
  if (f.gab) m |= JXLGPU_STAGE_GAB;
  if (f.epf_iters >= 3) m |= JXLGPU_STAGE_EPF0;
  if (f.epf_iters >= 1) m |= JXLGPU_STAGE_EPF1;
  if (f.epf_iters >= 2) m |= JXLGPU_STAGE_EPF2;
  return m;
}

template <uint32_t MASK>
void launch_strip_mask(jxlgpu_ctx* ctx, float* dev_out, size_t out_stride_floats, cudaStream_t s) {
  using C = StripCfg<MASK>;
  const FrameDev& P = ctx->P;
  const int band_h = (int)(P.band_y1 - P.band_y0);
  const int strips = ((int)P.xsize + C::kOutCols - 1) / C::kOutCols;
  // Exactly one wave: as many CTAs as fit on the chip at this kernel's occupancy (a 5% second
  // wave would double the kernel time), segments long enough to amortise the pipeline fill.
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, filter_strip_kernel<MASK>, kStripThreads,
                                                  C::kSmemBytes);
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  const int slots = ctx->num_sms * blocks_per_sm;
  int segs = slots / strips;
  if (segs < 1) segs = 1;
  int seg_rows = (band_h + segs - 1) / segs;
  if (seg_rows < 64) seg_rows = 64;
  seg_rows = (seg_rows + 7) & ~7;
  segs = (band_h + seg_rows - 1) / seg_rows;
  filter_strip_kernel<MASK><<<dim3(strips, segs), kStripThreads, C::kSmemBytes, s>>>(P, dev_out, out_stride_floats, seg_rows);
}

// the stage chains PreparePipeline can build for a VarDCT XYB frame (dec_cache.cc:151-170)
bool launch_strip(jxlgpu_ctx* ctx, float* dev_out, size_t out_stride_floats, cudaStream_t s) {
  if (ctx->force_generic_filter) return false;
  switch (ctx->P.stage_mask) {
    case 16: launch_strip_mask<16>(ctx, dev_out, out_stride_floats, s); return true;
    case 17: launch_strip_mask<17>(ctx, dev_out, out_stride_floats, s); return true;
    case 20: launch_strip_mask<20>(ctx, dev_out, out_stride_floats, s); return true;
    case 21: launch_strip_mask<21>(ctx, dev_out, out_stride_floats, s); return true;
    case 28: launch_strip_mask<28>(ctx, dev_out, out_stride_floats, s); return true;
    case 29: launch_strip_mask<29>(ctx, dev_out, out_stride_floats, s); return true;
    case 30: launch_strip_mask<30>(ctx, dev_out, out_stride_floats, s); return true;
    case 31: launch_strip_mask<31>(ctx, dev_out, out_stride_floats, s); return true;
    default: return false;
  }
}

template <uint32_t MASK>
cudaError_t strip_attr() {
  return cudaFuncSetAttribute(filter_strip_kernel<MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)StripCfg<MASK>::kSmemBytes);
}

int launch_all(jxlgpu_ctx* ctx, float* dev_out, size_t out_stride_floats, cudaStream_t s) {
  FrameDev& P = ctx->P;
  CU(cudaMemsetAsync(ctx->counts.p, 0, kNumStrategies * sizeof(uint32_t), s));
  const int want_sigma = (P.stage_mask & 14u) ? 1 : 0;
  const bool prof = ctx->profile;
  if (prof) CU(cudaEventRecord(ctx->prof_ev[0], s));
  plan_kernel<<<ctx->plan_groups, 1024, 0, s>>>(P, want_sigma);
  if (prof) CU(cudaEventRecord(ctx->prof_ev[1], s));
  const int grid8 = ctx->num_sms * 3;   // idct8_kernel: __launch_bounds__(256, 3)
  const int grid_mid = ctx->num_sms * 2;
  const int large_grid = ctx->num_sms * 2;
  if (P.ac_is32) idct8_kernel<true><<<grid8, kSmallWarpsPerCta * 32, 0, s>>>(P);
  else idct8_kernel<false><<<grid8, kSmallWarpsPerCta * 32, 0, s>>>(P);
  if (prof) CU(cudaEventRecord(ctx->prof_ev[2], s));
  if (P.ac_is32) idct_mid_kernel<true><<<grid_mid, kSmallWarpsPerCta * 32, 0, s>>>(P);
  else idct_mid_kernel<false><<<grid_mid, kSmallWarpsPerCta * 32, 0, s>>>(P);
  if (prof) CU(cudaEventRecord(ctx->prof_ev[3], s));
  if (P.ac_is32) idct_large_kernel<true><<<large_grid, 256, 0, s>>>(P);
  else idct_large_kernel<false><<<large_grid, 256, 0, s>>>(P);
  if (prof) CU(cudaEventRecord(ctx->prof_ev[4], s));
  const uint32_t band_h = P.band_y1 - P.band_y0;
  if (!launch_strip(ctx, dev_out, out_stride_floats, s)) {
    // stage chains outside the production set (test taps): generic tile kernel
    dim3 grid((P.xsize + kTW - 1) / kTW, (band_h + kTH - 1) / kTH);
    filter_kernel<<<grid, kFilterThreads, kFilterSmemFloats * sizeof(float), s>>>(P, dev_out, out_stride_floats);
  }
  if (prof) CU(cudaEventRecord(ctx->prof_ev[5], s));
  ctx->launches += 5;
  CU(cudaGetLastError());
  return JXLGPU_OK;
}

}  // namespace

extern "C" {

uint32_t jxlgpu_abi_version(void) { return JXLGPU_ABI_VERSION; }

const char* jxlgpu_error_string(int code) {
  switch (code) {
    case JXLGPU_OK: return "ok";
    case JXLGPU_ERR_INVALID_ARGUMENT: return "invalid argument";
    case JXLGPU_ERR_UNSUPPORTED: return "frame not eligible for the GPU path";
    case JXLGPU_ERR_NO_DEVICE: return "no CUDA device";
    case JXLGPU_ERR_CUDA: return "CUDA error";
    case JXLGPU_ERR_OUT_OF_MEMORY: return "out of device memory";
    case JXLGPU_ERR_STATE: return "call out of order";
  }
  return "unknown";
}

const char* jxlgpu_last_error(const jxlgpu_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int jxlgpu_create(jxlgpu_ctx** out, const jxlgpu_config* cfg) {
  if (!out || !cfg || cfg->abi_version != JXLGPU_ABI_VERSION) return JXLGPU_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0 || cfg->device < 0 || cfg->device >= n) return JXLGPU_ERR_NO_DEVICE;
  jxlgpu_ctx* ctx = new jxlgpu_ctx();
  ctx->device = cfg->device;
  ctx->num_threads = cfg->num_host_threads ? cfg->num_host_threads : 1;
  auto bail = [&](cudaError_t err, const char* what) {
    fprintf(stderr, "jxlgpu_create: %s: %s\n", what, cudaGetErrorString(err));
    delete ctx;
    return JXLGPU_ERR_CUDA;
  };
  if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, ctx->device)) != cudaSuccess) return bail(e, "props");
  ctx->num_sms = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
  ctx->up_streams.resize(ctx->num_threads);
  ctx->up_events.resize(ctx->num_threads);
  for (uint32_t i = 0; i < ctx->num_threads; i++) {
    if ((e = cudaStreamCreateWithFlags(&ctx->up_streams[i], cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    if ((e = cudaEventCreateWithFlags(&ctx->up_events[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
  }
  if ((e = cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kFilterSmemFloats * sizeof(float)))) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(filter_kernel)");
  for (cudaError_t ea : {strip_attr<16>(), strip_attr<17>(), strip_attr<20>(), strip_attr<21>(), strip_attr<28>(),
                         strip_attr<29>(), strip_attr<30>(), strip_attr<31>()})
    if (ea != cudaSuccess) return bail(ea, "cudaFuncSetAttribute(filter_strip_kernel)");
  {
    const char* env = getenv("JXLGPU_FORCE_GENERIC_FILTER");
    ctx->force_generic_filter = env && env[0] == '1';
  }
  if ((e = ctx->counts.ensure(kNumStrategies * sizeof(uint32_t))) != cudaSuccess) return bail(e, "alloc");
  for (auto& ev : ctx->prof_ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "event");
  *out = ctx;
  return JXLGPU_OK;
}

void jxlgpu_destroy(jxlgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (DevBuf* b : {&ctx->acs, &ctx->quant, &ctx->sharp, &ctx->ytox, &ctx->ytob, &ctx->dc, &ctx->dq,
                    &ctx->coeff[0], &ctx->coeff[1], &ctx->coeff[2], &ctx->coeff_off, &ctx->sigma,
                    &ctx->list, &ctx->counts, &ctx->xyb, &ctx->out})
    b->release();
  for (auto s : ctx->up_streams) cudaStreamDestroy(s);
  for (auto ev : ctx->up_events) cudaEventDestroy(ev);
  for (auto ev : ctx->prof_ev) if (ev) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int jxlgpu_frame_begin(jxlgpu_ctx* ctx, const jxlgpu_frame* f) {
  if (!ctx || !f) return JXLGPU_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  if (f->xsize == 0 || f->ysize == 0 || f->xsize_blocks != (f->xsize + 7) / 8 ||
      f->ysize_blocks != (f->ysize + 7) / 8 || f->xsize_blocks > 65535 || f->ysize_blocks > 65535)
    return JXLGPU_ERR_INVALID_ARGUMENT;
  if (!f->ac_strategy || !f->raw_quant || !f->ytox_map || !f->ytob_map || !f->dc[0] || !f->dc[1] ||
      !f->dc[2] || !f->dequant_table)
    return JXLGPU_ERR_INVALID_ARGUMENT;
  if (f->ac_type > JXLGPU_AC_INT32 || f->out_format > JXLGPU_OUT_PLANAR_F32) return JXLGPU_ERR_INVALID_ARGUMENT;
  const uint32_t mask = effective_mask(*f);
  if ((mask & 14u) && !f->epf_sharpness) return JXLGPU_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < 3 * kNumStrategies; i++) {
    const size_t n = (size_t)64 * covered_x(i / 3) * covered_y(i / 3);
    if (f->dequant_offsets[i] + n > f->dequant_table_floats) return JXLGPU_ERR_INVALID_ARGUMENT;
    if (f->dequant_offsets[i] % 4) return JXLGPU_ERR_INVALID_ARGUMENT;  // 16-byte vector loads
  }
  FrameDev& P = ctx->P;
  const size_t xb = f->xsize_blocks, yb = f->ysize_blocks, nblocks = xb * yb;
  P.xsize = f->xsize; P.ysize = f->ysize; P.xb = xb; P.yb = yb;
  P.xg = (xb + 31) / 32; P.yg = (yb + 31) / 32;
  ctx->num_groups = P.xg * P.yg;
  P.ac_is32 = f->ac_type == JXLGPU_AC_INT32;
  ctx->elem_size = P.ac_is32 ? 4 : 2;
  P.stage_mask = mask;
  P.out_format = f->out_format;
  if (f->band_ny_groups == 0) {
    P.band_y0 = 0; P.band_y1 = f->ysize;
  } else {
    if (f->band_y0_groups + f->band_ny_groups > P.yg) return JXLGPU_ERR_INVALID_ARGUMENT;
    P.band_y0 = f->band_y0_groups * 256u;
    const uint32_t y1 = (f->band_y0_groups + f->band_ny_groups) * 256u;
    P.band_y1 = y1 < f->ysize ? y1 : f->ysize;
  }
  {
    const uint32_t halo = ((mask & 1) ? 1 : 0) + ((mask & 2) ? 3 : 0) + ((mask & 4) ? 2 : 0) + ((mask & 8) ? 1 : 0);
    P.need_y0 = P.band_y0 > halo ? P.band_y0 - halo : 0;
    P.need_y1 = P.band_y1 + halo < f->ysize ? P.band_y1 + halo : f->ysize;
    if (P.band_y1 >= f->ysize) P.need_y1 = (uint32_t)yb * 8;  // bottom band also owns the padded block rows
    const uint32_t gy0 = P.need_y0 / 256, gy1 = (P.need_y1 + 255) / 256;
    ctx->need_g0 = gy0 * P.xg;
    ctx->need_g1 = (gy1 < P.yg ? gy1 : P.yg) * P.xg;
    P.plan_g0 = ctx->need_g0;
    ctx->plan_groups = ctx->need_g1 - ctx->need_g0;
  }
  const size_t cmw = (xb + 7) / 8, cmh = (yb + 7) / 8;
  CU(ctx->acs.ensure(nblocks));
  CU(ctx->quant.ensure(nblocks * 4));
  CU(ctx->sharp.ensure(nblocks));
  CU(ctx->ytox.ensure(cmw * cmh));
  CU(ctx->ytob.ensure(cmw * cmh));
  CU(ctx->dc.ensure(3 * nblocks * 4));
  CU(ctx->dq.ensure(f->dequant_table_floats * 4));
  CU(ctx->coeff_off.ensure(nblocks * 2));
  CU(ctx->sigma.ensure(nblocks * 4));
  // per-strategy work lists, capacity = max number of varblocks of that size
  size_t total = 0;
  for (int s = 0; s < kNumStrategies; s++) {
    P.list_base[s] = (uint32_t)total;
    total += nblocks / (covered_x(s) * covered_y(s)) + 1;
  }
  CU(ctx->list.ensure(total * 4));
  P.row_stride = xb * 8;
  P.plane_stride = P.row_stride * yb * 8;
  CU(ctx->xyb.ensure(3 * P.plane_stride * 4));
  if (!ctx->coeff_external)
    for (int c = 0; c < 3; c++) CU(ctx->coeff[c].ensure((size_t)ctx->num_groups * 65536 * ctx->elem_size));
  cudaStream_t s = ctx->stream;
  CU(upload_plane<uint8_t>(ctx->acs.p, f->ac_strategy, f->ac_strategy_stride, xb, yb, s));
  CU(upload_plane<int32_t>(ctx->quant.p, f->raw_quant, f->raw_quant_stride, xb, yb, s));
  if (f->epf_sharpness) CU(upload_plane<uint8_t>(ctx->sharp.p, f->epf_sharpness, f->epf_sharpness_stride, xb, yb, s));
  CU(upload_plane<int8_t>(ctx->ytox.p, f->ytox_map, f->cmap_stride, cmw, cmh, s));
  CU(upload_plane<int8_t>(ctx->ytob.p, f->ytob_map, f->cmap_stride, cmw, cmh, s));
  for (int c = 0; c < 3; c++)
    CU(upload_plane<float>((float*)ctx->dc.p + c * nblocks, f->dc[c], f->dc_stride, xb, yb, s));
  CU(cudaMemcpyAsync(ctx->dq.p, f->dequant_table, f->dequant_table_floats * 4, cudaMemcpyHostToDevice, s));
  P.acs = (const uint8_t*)ctx->acs.p;
  P.quant = (const int32_t*)ctx->quant.p;
  P.sharp = (const uint8_t*)ctx->sharp.p;
  P.ytox = (const int8_t*)ctx->ytox.p;
  P.ytob = (const int8_t*)ctx->ytob.p;
  P.cmap_stride = cmw;
  P.dc = (const float*)ctx->dc.p;
  P.dq = (const float*)ctx->dq.p;
  memcpy(P.dq_off, f->dequant_offsets, sizeof(P.dq_off));
  if (!ctx->coeff_external)
    for (int c = 0; c < 3; c++) P.coeff[c] = ctx->coeff[c].p;
  P.coeff_off = (uint16_t*)ctx->coeff_off.p;
  P.sigma = (float*)ctx->sigma.p;
  P.list = (uint32_t*)ctx->list.p;
  P.counts = (uint32_t*)ctx->counts.p;
  P.xyb = (float*)ctx->xyb.p;
  P.inv_global_scale = f->inv_global_scale;
  P.quant_scale = f->quant_scale;
  P.x_dm = f->x_dm_multiplier;
  P.b_dm = f->b_dm_multiplier;
  memcpy(P.qbias, f->quant_biases, sizeof(P.qbias));
  P.cfl_base_x = f->cfl_base_x; P.cfl_base_b = f->cfl_base_b; P.cfl_scale = f->cfl_color_scale;
  for (int c = 0; c < 3; c++) {  // GaborishStage ctor (stage_gaborish.cc:33-54)
    float w0 = 1.0f, w1 = f->gab_weights[2 * c], w2 = f->gab_weights[2 * c + 1];
    const float div = w0 + 4 * (w1 + w2);
    const float mul = 1.0f / div;
    P.gab_w[3 * c] = w0 * mul; P.gab_w[3 * c + 1] = w1 * mul; P.gab_w[3 * c + 2] = w2 * mul;
  }
  memcpy(P.epf_sharp_lut, f->epf_sharp_lut, sizeof(P.epf_sharp_lut));
  memcpy(P.epf_scale, f->epf_channel_scale, sizeof(P.epf_scale));
  P.epf_quant_mul = f->epf_quant_mul;
  P.epf_sm[0] = (float)(f->epf_pass0_sigma_scale * 1.65);  // stage_epf.cc:93
  P.epf_sm[1] = 1.65f;                                      // stage_epf.cc:236
  P.epf_sm[2] = (float)(f->epf_pass2_sigma_scale * 1.65);  // stage_epf.cc:427
  P.epf_border_mul = f->epf_border_sad_mul;
  memcpy(P.opsin_m, f->inverse_opsin_matrix, sizeof(P.opsin_m));
  memcpy(P.opsin_bias, f->opsin_biases, sizeof(P.opsin_bias));
  memcpy(P.opsin_cbrt, f->opsin_biases_cbrt, sizeof(P.opsin_cbrt));
  ctx->out_stride_floats = out_floats_per_row(*f);
  ctx->submitted.assign(ctx->num_groups, ctx->coeff_external ? 1 : 0);
  ctx->in_frame = true;
  return JXLGPU_OK;
}

int jxlgpu_submit_group(jxlgpu_ctx* ctx, uint32_t g, size_t thread_id, const void* const coeff[3], size_t ncoeff) {
  if (!ctx || !coeff) return JXLGPU_ERR_INVALID_ARGUMENT;
  if (!ctx->in_frame || ctx->coeff_external) return JXLGPU_ERR_STATE;
  if (g >= ctx->num_groups || thread_id >= ctx->num_threads || ncoeff > 65536) return JXLGPU_ERR_INVALID_ARGUMENT;
  // NB: cudaSetDevice is per host thread
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return JXLGPU_ERR_CUDA;
  cudaStream_t s = ctx->up_streams[thread_id];
  for (int c = 0; c < 3; c++) {
    if (!coeff[c]) return JXLGPU_ERR_INVALID_ARGUMENT;
    e = cudaMemcpyAsync((uint8_t*)ctx->coeff[c].p + (size_t)g * 65536 * ctx->elem_size, coeff[c],
                        ncoeff * ctx->elem_size, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) {
      std::lock_guard<std::mutex> lk(ctx->mu);
      return fail_cuda(ctx, e, "cudaMemcpyAsync(coefficients)");
    }
  }
  ctx->submitted[g] = 1;
  return JXLGPU_OK;
}

int jxlgpu_set_device_coefficients(jxlgpu_ctx* ctx, const void* const dev_coeff[3]) {
  if (!ctx) return JXLGPU_ERR_INVALID_ARGUMENT;
  if (!dev_coeff) {
    ctx->coeff_external = false;
    return JXLGPU_OK;
  }
  ctx->coeff_external = true;
  for (int c = 0; c < 3; c++) ctx->P.coeff[c] = dev_coeff[c];
  if (ctx->in_frame) ctx->submitted.assign(ctx->num_groups, 1);
  return JXLGPU_OK;
}

int jxlgpu_render_device(jxlgpu_ctx* ctx, void* dev_out, size_t out_stride_bytes, void* cuda_stream) {
  if (!ctx) return JXLGPU_ERR_INVALID_ARGUMENT;
  if (!ctx->in_frame) return JXLGPU_ERR_STATE;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
  if (cuda_stream) {  // side info was uploaded on the context stream
    CU(cudaEventRecord(ctx->up_events[0], ctx->stream));
    CU(cudaStreamWaitEvent(s, ctx->up_events[0], 0));
  }
  const uint32_t band_h = ctx->P.band_y1 - ctx->P.band_y0;
  float* o = (float*)dev_out;
  size_t stride = out_stride_bytes / 4;
  if (!o) {
    const size_t planes = ctx->P.out_format == JXLGPU_OUT_RGB_F32 ? 1 : 3;
    CU(ctx->out.ensure(planes * band_h * ctx->out_stride_floats * 4));
    o = (float*)ctx->out.p;
    stride = ctx->out_stride_floats;
  } else if (out_stride_bytes % 4 || stride < ctx->out_stride_floats) {
    return JXLGPU_ERR_INVALID_ARGUMENT;
  }
  return launch_all(ctx, o, stride, s);
}

int jxlgpu_frame_finish(jxlgpu_ctx* ctx, void* out, size_t out_stride_bytes) {
  if (!ctx) return JXLGPU_ERR_INVALID_ARGUMENT;
  if (!ctx->in_frame) return JXLGPU_ERR_STATE;
  CU(cudaSetDevice(ctx->device));
  // every group of the band (+ halo rows) must have arrived
  for (uint32_t g = ctx->need_g0; g < ctx->need_g1; g++)
    if (!ctx->submitted[g]) { ctx->last_error = "missing group"; return JXLGPU_ERR_STATE; }
  for (uint32_t i = 0; i < ctx->num_threads; i++) {
    CU(cudaEventRecord(ctx->up_events[i], ctx->up_streams[i]));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->up_events[i], 0));
  }
  int rc = jxlgpu_render_device(ctx, nullptr, 0, nullptr);
  if (rc) return rc;
  if (out) {
    const uint32_t band_h = ctx->P.band_y1 - ctx->P.band_y0;
    const size_t row_bytes = ctx->out_stride_floats * 4;
    if (out_stride_bytes < row_bytes) return JXLGPU_ERR_INVALID_ARGUMENT;
    const size_t planes = ctx->P.out_format == JXLGPU_OUT_RGB_F32 ? 1 : 3;
    CU(cudaMemcpy2DAsync(out, out_stride_bytes, ctx->out.p, row_bytes, row_bytes, planes * band_h,
                         cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->in_frame = false;
  return JXLGPU_OK;
}

int jxlgpu_device_output(jxlgpu_ctx* ctx, void** dev_ptr, size_t* stride_bytes) {
  if (!ctx || !dev_ptr || !stride_bytes) return JXLGPU_ERR_INVALID_ARGUMENT;
  *dev_ptr = ctx->out.p;
  *stride_bytes = ctx->out_stride_floats * 4;
  return ctx->out.p ? JXLGPU_OK : JXLGPU_ERR_STATE;
}

int jxlgpu_device_xyb(jxlgpu_ctx* ctx, float** dev_ptr, size_t* plane_stride_floats, size_t* row_stride_floats) {
  if (!ctx || !dev_ptr) return JXLGPU_ERR_INVALID_ARGUMENT;
  *dev_ptr = (float*)ctx->xyb.p;
  if (plane_stride_floats) *plane_stride_floats = ctx->P.plane_stride;
  if (row_stride_floats) *row_stride_floats = ctx->P.row_stride;
  return ctx->xyb.p ? JXLGPU_OK : JXLGPU_ERR_STATE;
}

int jxlgpu_synchronize(jxlgpu_ctx* ctx) {
  if (!ctx) return JXLGPU_ERR_INVALID_ARGUMENT;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return JXLGPU_OK;
}

uint64_t jxlgpu_launch_count(const jxlgpu_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int jxlgpu_set_profiling(jxlgpu_ctx* ctx, int enable) {
  if (!ctx) return JXLGPU_ERR_INVALID_ARGUMENT;
  ctx->profile = enable != 0;
  return JXLGPU_OK;
}

int jxlgpu_kernel_times(jxlgpu_ctx* ctx, float ms[5]) {
  if (!ctx || !ms) return JXLGPU_ERR_INVALID_ARGUMENT;
  if (!ctx->profile) return JXLGPU_ERR_STATE;
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventSynchronize(ctx->prof_ev[5]));
  for (int i = 0; i < 5; i++) CU(cudaEventElapsedTime(&ms[i], ctx->prof_ev[i], ctx->prof_ev[i + 1]));
  return JXLGPU_OK;
}

void* jxlgpu_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}

void jxlgpu_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
