// libjxl_gpu_backend.h -- the libjxl-side half of the drop-in: what FrameDecoder calls at the three sites
// INTEGRATION.md §2 names.  This header is OURS (it only includes reference headers and include/jxl_b200.h);
// integration/patch_libjxl.py inserts the one-line calls into a BUILD-TIME COPY of lib/jxl/dec_frame.cc and
// lib/jxl/dec_group.cc (never committed; oracle/_ref/patched/), and oracle/build_ref.py links that variant
// ("gpu") against libjxl_b200.so.  The public JxlDecoder API and the JxlParallelRunner stay untouched: an
// application keeps calling JxlDecoderSetParallelRunner / JxlDecoderSetImageOutBuffer / JxlDecoderProcessInput.
//
//   site 1  FrameDecoder::ProcessACGlobal (lib/jxl/dec_frame.cc:417-434)   -> WantFrame(): accumulate-mode
//           coefficient storage in page-locked, group-major memory (pinned_ac_image.h)
//   site 2  FrameDecoder::ProcessSections, after ProcessACGlobal (:693-696) -> BeginFrame(): jxlgpu_frame_begin
//           with pointers into the decoder's own images (gpu_frame_binding.h) + jxlgpu_frame_set_output
//   site 3  FrameDecoder::ProcessACGroup (:483-560): DecodeGroup runs with draw == kDontDraw
//           (lib/jxl/dec_group.cc:357-363,724-727 -> DontDraw()), then GroupDecoded(): the thread that
//           completes a row of AC groups hands the whole row to jxlgpu_submit_groups (one DMA)
//   site 4  FrameDecoder::FinalizeFrame (:860-882)                          -> FinishFrame(): jxlgpu_frame_finish
//
// A frame that is not eligible, a process without a CUDA device, or JXLB_GPU_BACKEND=0 keeps libjxl's CPU
// path: every hook then returns "not mine".  Errors after a frame was taken are decode errors (jxl::Status).
#ifndef JXL_B200_INTEGRATION_LIBJXL_GPU_BACKEND_H_
#define JXL_B200_INTEGRATION_LIBJXL_GPU_BACKEND_H_

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "gpu_frame_binding.h"
#include "jxl_b200.h"
#include "lib/jxl/base/status.h"
#include "lib/jxl/blending.h"
#include "lib/jxl/dec_cache.h"
#include "lib/jxl/frame_header.h"
#include "lib/jxl/image_bundle.h"
#include "lib/jxl/render_pipeline/stage_tone_mapping.h"
#include "pinned_ac_image.h"

namespace jxlb_integration {

// dec_group.cc asks this for every group (defined in the dec_frame.cc translation unit)
bool DontDraw(const jxl::PassesDecoderState* dec_state);

// Sparse hand-off: while a worker thread entropy-decodes one AC group of a GPU frame, the
// non-zero coefficients are appended to this per-thread sink instead of being added into a zero-filled
// block (lib/jxl/dec_group.cc:515-534): (pos << 16) | u16 value words, or {pos, value} pairs for the rare
// values outside int16.  pos = offset of the varblock inside the group + coefficient index.
struct SparseSink {
  uint32_t* nz16[3];
  uint32_t* nz32[3];
  uint32_t n16[3], n32[3];
  uint32_t offset;  // of the varblock being decoded (set where DecodeGroupImpl computes qblock[])
  inline void Append(size_t c, uint32_t k, ptrdiff_t coeff) {
    const uint32_t pos = offset + k;
    if (coeff >= -32768 && coeff < 32768) {
      nz16[c][n16[c]++] = (pos << 16) | (static_cast<uint32_t>(coeff) & 0xffffu);
    } else {
      uint32_t* p = nz32[c] + 2 * static_cast<size_t>(n32[c]++);
      p[0] = pos;
      p[1] = static_cast<uint32_t>(static_cast<int32_t>(coeff));
    }
  }
};
SparseSink* CurrentSink();            // null on the CPU path and for dense GPU frames

#ifdef JXLB_GPU_BACKEND_IMPLEMENTATION

struct GpuFrame {
  GpuFrameBinding binding;
  std::vector<std::atomic<uint32_t>> row_count;  // decoded groups per AC-group row (dense hand-off)
  uint32_t xg = 0, yg = 0;
  bool is16 = false;
  bool sparse = false;
  bool begun = false;
  std::atomic<size_t> arena_words{0};            // sparse: bump allocation inside the pinned arena
  std::atomic<int> error{0};
};

// One page-locked arena per process, grown on demand and reused by every frame (cudaHostAlloc of hundreds
// of megabytes costs more than decoding the frame).  One GPU frame is in flight at a time (the C ABI's
// contract), so the arena has one user.
struct PinnedArena {
  void* p = nullptr;
  size_t cap = 0;
  void* Get(size_t bytes) {
    if (bytes > cap) {
      if (p) jxlgpu_free_pinned(p);
      p = jxlgpu_alloc_pinned(bytes);
      cap = p ? bytes : 0;
    }
    return p;
  }
};

struct GpuBackend {
  std::mutex mu;
  jxlgpu_ctx* ctx = nullptr;
  bool tried = false;
  bool want_sparse = true;
  PinnedArena arena;
  std::unordered_map<const void*, std::unique_ptr<GpuFrame>> frames;  // key: PassesDecoderState*
  uint64_t frames_taken = 0;

  static GpuBackend& Get() {
    static GpuBackend b;
    return b;
  }
  // one context per process, created on the first eligible frame; any failure = CPU path for good
  jxlgpu_ctx* Context() {
    std::lock_guard<std::mutex> lk(mu);
    if (!tried) {
      tried = true;
      const char* e = getenv("JXLB_GPU_BACKEND");
      if (!(e && e[0] == '0')) {
        jxlgpu_config cfg = {JXLGPU_ABI_VERSION, 0, 1, 0};
        if (const char* d = getenv("JXLB_GPU_DEVICE")) cfg.device = atoi(d);
        if (jxlgpu_create(&ctx, &cfg) != JXLGPU_OK) ctx = nullptr;
      }
    }
    return ctx;
  }
  GpuFrame* Find(const void* dec_state) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = frames.find(dec_state);
    return it == frames.end() ? nullptr : it->second.get();
  }
};

inline uint64_t FramesTaken() { return GpuBackend::Get().frames_taken; }

// coefficient storage views over the arena (the arena outlives them: nothing to free)
inline void* ArenaAlloc(size_t bytes) { return GpuBackend::Get().arena.Get(bytes); }
inline void ArenaFree(void*) {}

// accumulate-mode storage for a sparse frame: DecodeGroupImpl only needs PlaneRow() to hand *something* to
// LoadBlock (lib/jxl/dec_group.cc:335-338); the patched DecodeACVarBlock never dereferences it
class SinkACImage final : public jxl::ACImage {
 public:
  explicit SinkACImage(bool is16) : is16_(is16) {}
  jxl::ACType Type() const override { return is16_ ? jxl::ACType::k16 : jxl::ACType::k32; }
  jxl::ACPtr PlaneRow(size_t, size_t, size_t) override {
    jxl::ACPtr p;
    if (is16_) p.ptr16 = reinterpret_cast<int16_t*>(dummy_);
    else p.ptr32 = reinterpret_cast<int32_t*>(dummy_);
    return p;
  }
  jxl::ConstACPtr PlaneRow(size_t, size_t, size_t) const override {
    jxl::ConstACPtr p;
    if (is16_) p.ptr16 = reinterpret_cast<const int16_t*>(dummy_);
    else p.ptr32 = reinterpret_cast<const int32_t*>(dummy_);
    return p;
  }
  size_t PixelsPerRow() const override { return 0; }
  void ZeroFill() override {}
  void ZeroFillPlane(size_t) override {}
  bool IsEmpty() const override { return false; }

 private:
  bool is16_;
  alignas(16) uint8_t dummy_[64] = {};
};

// per worker thread: worst-case list space for ONE group (3 channels x 65536 words + as many pairs)
struct ThreadSink {
  SparseSink sink;
  std::unique_ptr<uint32_t[]> buf;
  bool active = false;
  void Arm() {
    constexpr size_t kG = 65536;
    if (!buf) buf.reset(new uint32_t[3 * kG + 6 * kG]);
    for (int c = 0; c < 3; c++) {
      sink.nz16[c] = buf.get() + c * kG;
      sink.nz32[c] = buf.get() + 3 * kG + c * 2 * kG;
      sink.n16[c] = sink.n32[c] = 0;
    }
    sink.offset = 0;
    active = true;
  }
};
inline ThreadSink& TlsSink() {
  static thread_local ThreadSink t;
  return t;
}
SparseSink* CurrentSink() {
  ThreadSink& t = TlsSink();
  return t.active ? &t.sink : nullptr;
}

// JxlPixelFormat of the application's buffer -> JXLGPU_OUT_* (INTEGRATION.md §1); false: CPU path
inline bool MapOutput(const jxl::PassesDecoderState& ds, bool has_alpha, uint32_t* out_format) {
  const jxl::ImageOutput& o = ds.main_output;
  if (!o.buffer || o.callback.IsPresent() || ds.undo_orientation != jxl::Orientation::kIdentity || ds.unpremul_alpha ||
      ds.fast_xyb_srgb8_conversion || has_alpha)
    return false;
  const uint16_t one = 1;
  const bool little = *reinterpret_cast<const uint8_t*>(&one) == 1;
  const bool native = o.format.endianness == JXL_NATIVE_ENDIAN ||
                      o.format.endianness == (little ? JXL_LITTLE_ENDIAN : JXL_BIG_ENDIAN);
  switch (o.format.data_type) {
    case JXL_TYPE_FLOAT:
      if (o.format.num_channels != 3 || !native) return false;
      *out_format = JXLGPU_OUT_RGB_F32;
      return true;
    case JXL_TYPE_UINT8:
      if (o.bits_per_sample != 8) return false;
      if (o.format.num_channels == 3) *out_format = JXLGPU_OUT_RGB_U8;
      else if (o.format.num_channels == 4) *out_format = JXLGPU_OUT_RGBA_U8;
      else return false;
      return true;
    case JXL_TYPE_UINT16:
      if (o.format.num_channels != 3 || !native || o.bits_per_sample != 16) return false;
      *out_format = JXLGPU_OUT_RGB_U16;
      return true;
    case JXL_TYPE_FLOAT16:
      if (o.format.num_channels != 3 || !native) return false;
      *out_format = JXLGPU_OUT_RGB_F16;
      return true;
    default:
      return false;
  }
}

// The frames whose render pipeline is exactly [Gaborish] [EPF..] XYB [FromLinear(sRGB)] WriteToOutput
// (PassesDecoderState::PreparePipeline, lib/jxl/dec_cache.cc:117-345) into a caller-owned buffer.
inline bool FrameIsOurs(const jxl::FrameHeader& fh, const jxl::PassesDecoderState& ds, const jxl::ImageBundle* decoded,
                        uint32_t* out_format, uint32_t* stage_mask) {
  const jxl::CodecMetadata& md = *fh.nonserialized_metadata;
  if (!IsEligible(fh, md) || decoded->IsJPEG()) return false;
  if (fh.frame_type != jxl::FrameType::kRegularFrame || fh.dc_level != 0 || fh.CanBeReferenced() ||
      jxl::NeedsBlending(fh) || fh.custom_size_or_origin || fh.nonserialized_is_preview)
    return false;
  const auto& oei = ds.output_encoding_info;
  const auto& ce = oei.color_encoding;
  if (ce.GetColorSpace() != jxl::ColorSpace::kRGB) return false;
  if (!(oei.color_encoding_is_original || !oei.cms_set)) return false;   // a CMS stage would follow
  if (jxl::GetToneMappingStage(oei)) return false;
  if (fh.color_transform == jxl::ColorTransform::kYCbCr) {
    // a non-XYB image: kYCbCrStage leaves the pixels in the image's own encoding, nothing follows it unless the
    // application asked for another colour space (then a CMS stage would: dec_cache.cc:355-360)
    if (!oei.color_encoding_is_original) return false;
    *stage_mask = 0;
  } else if (ce.Tf().IsLinear()) *stage_mask = 0;
  else if (ce.Tf().IsSRGB()) *stage_mask = JXLGPU_STAGE_SRGB;
  else return false;
  if (ds.width != ds.shared->frame_dim.xsize_upsampled || ds.height != ds.shared->frame_dim.ysize_upsampled) return false;
  return MapOutput(ds, /*has_alpha=*/false, out_format);
}

// ---- site 1: coefficient storage (ProcessACGlobal).  `use_16_bit` as the reference computed it. ----
// Returns true and installs pinned group-major accumulate-mode storage when the frame will go to the GPU.
inline bool WantFrame(const jxl::FrameHeader& fh, jxl::PassesDecoderState* ds, const jxl::ImageBundle* decoded,
                      bool use_16_bit, size_t num_groups) {
  uint32_t fmt = 0, mask = 0;
  if (!FrameIsOurs(fh, *ds, decoded, &fmt, &mask)) return false;
  GpuBackend& be = GpuBackend::Get();
  if (!be.Context()) return false;
  // page-locked staging: dense = the [group][channel][65536] blocks themselves; sparse = the groups' non-zero
  // lists, bump-allocated (a group never needs more words than its dense block: larger lists are handed
  // over dense, see GroupDecoded)
  std::unique_ptr<jxl::ACImage> store;
  // JXLB_GPU_SPARSE=0: dense [group][3][65536] blocks instead of non-zero lists (read per frame)
  const char* sp = getenv("JXLB_GPU_SPARSE");
  // a multi-pass (progressive) frame accumulates its passes in the dense blocks; the lists hold each position once
  const bool sparse = (sp ? sp[0] != '0' : be.want_sparse) && fh.passes.num_passes == 1;
  if (sparse) {
    if (!ArenaAlloc(num_groups * 3 * 65536 * sizeof(uint32_t))) return false;
    store.reset(new SinkACImage(use_16_bit));
  } else {
    // (no frame-wide ZeroFill: every worker zero-fills its group right before decoding it, PrepareGroup)
    if (use_16_bit) store = GroupMajorACImage<int16_t>::Make(num_groups, ArenaAlloc, ArenaFree);
    else store = GroupMajorACImage<int32_t>::Make(num_groups, ArenaAlloc, ArenaFree);
  }
  if (!store) return false;
  ds->coefficients = std::move(store);
  auto fr = std::make_unique<GpuFrame>();
  fr->is16 = use_16_bit;
  fr->sparse = sparse;
  fr->binding.frame.out_format = fmt;
  fr->binding.frame.stage_mask = mask;
  std::lock_guard<std::mutex> lk(be.mu);
  be.frames[ds] = std::move(fr);
  return true;
}

// ---- site 2: after ProcessACGlobal, DC finalised, output buffer known ----
inline jxl::Status BeginFrame(const jxl::FrameHeader& fh, jxl::PassesDecoderState* ds) {
  GpuBackend& be = GpuBackend::Get();
  GpuFrame* fr = be.Find(ds);
  if (!fr || fr->begun) return true;
  const uint32_t fmt = fr->binding.frame.out_format, mask = fr->binding.frame.stage_mask;
  if (!BindGpuFrame(*ds, fh, fmt, mask, &fr->binding)) return JXL_FAILURE("gpu backend: frame binding failed");
  const jxl::FrameDimensions& d = ds->shared->frame_dim;
  fr->xg = static_cast<uint32_t>(d.xsize_groups);
  fr->yg = static_cast<uint32_t>(d.ysize_groups);
  fr->row_count = std::vector<std::atomic<uint32_t>>(fr->yg);
  for (auto& c : fr->row_count) c.store(0);
  if (jxlgpu_frame_begin(be.ctx, &fr->binding.frame) != JXLGPU_OK)
    return JXL_FAILURE("gpu backend: frame_begin: %s", jxlgpu_last_error(be.ctx));
  if (jxlgpu_frame_set_output(be.ctx, ds->main_output.buffer, ds->main_output.stride) != JXLGPU_OK)
    return JXL_FAILURE("gpu backend: frame_set_output");
  fr->begun = true;
  return true;
}

// ---- site 3: one AC group has been entropy-decoded into the pinned storage ----
inline bool FrameActive(const jxl::PassesDecoderState* ds) {
  GpuFrame* fr = GpuBackend::Get().Find(ds);
  return fr != nullptr;
}
// before DecodeGroup: dense -> zero-fill this group's blocks (accumulate mode adds into them,
// lib/jxl/dec_group.cc:527-531); sparse -> arm the calling thread's sink
inline void PrepareGroup(jxl::PassesDecoderState* ds, size_t group, size_t passes_done) {
  GpuFrame* fr = GpuBackend::Get().Find(ds);
  if (!fr) return;
  if (fr->sparse) {
    TlsSink().Arm();
    return;
  }
  if (passes_done != 0) return;  // later passes add to what the earlier ones left (dec_group.cc:527-531)
  const jxl::Rect br = ds->shared->frame_dim.BlockGroupRect(group);
  const size_t n = 64 * br.xsize() * br.ysize();
  for (size_t c = 0; c < 3; c++) {
    jxl::ACPtr p = ds->coefficients->PlaneRow(c, group, 0);
    if (fr->is16) memset(p.ptr16, 0, n * sizeof(int16_t));
    else memset(p.ptr32, 0, n * sizeof(int32_t));
  }
}

inline jxl::Status GroupDecodedSparse(GpuBackend& be, GpuFrame* fr, jxl::PassesDecoderState* ds, size_t group) {
  ThreadSink& t = TlsSink();
  t.active = false;
  const SparseSink& k = t.sink;
  size_t words = 0;
  for (int c = 0; c < 3; c++) words += k.n16[c] + 2 * static_cast<size_t>(k.n32[c]);
  words += 1;  // pair lists start on an 8-byte boundary
  const size_t kDense = 3 * 65536;
  uint32_t* arena = static_cast<uint32_t*>(be.arena.p);
  if (words > kDense) {
    // more list words than the dense block has coefficients (a pathological frame): expand on the host into
    // this group's share of the arena and hand it over dense
    const size_t at = fr->arena_words.fetch_add(kDense);
    const size_t es = fr->is16 ? 2 : 4;
    uint8_t* blk = reinterpret_cast<uint8_t*>(arena + at);
    memset(blk, 0, kDense * es);
    for (int c = 0; c < 3; c++) {
      for (uint32_t i = 0; i < k.n16[c]; i++) {
        const uint32_t w = k.nz16[c][i];
        const int32_t v = static_cast<int16_t>(w & 0xffffu);
        if (fr->is16) reinterpret_cast<int16_t*>(blk)[c * 65536 + (w >> 16)] = static_cast<int16_t>(v);
        else reinterpret_cast<int32_t*>(blk)[c * 65536 + (w >> 16)] = v;
      }
      for (uint32_t i = 0; i < k.n32[c]; i++)
        reinterpret_cast<int32_t*>(blk)[c * 65536 + k.nz32[c][2 * i]] = static_cast<int32_t>(k.nz32[c][2 * i + 1]);
    }
    const void* co[3] = {blk, blk + 65536 * es, blk + 2 * 65536 * es};
    const jxl::Rect br = ds->shared->frame_dim.BlockGroupRect(group);
    if (jxlgpu_submit_group(be.ctx, static_cast<uint32_t>(group), 0, co, 64 * br.xsize() * br.ysize()) != JXLGPU_OK)
      return JXL_FAILURE("gpu backend: submit_group: %s", jxlgpu_last_error(be.ctx));
    return true;
  }
  // compact the three channels' lists next to each other in the arena: one DMA per kind of list
  size_t at = fr->arena_words.fetch_add(words);
  jxlgpu_sparse_group g = {};
  g.group_idx = static_cast<uint32_t>(group);
  for (int c = 0; c < 3; c++) {
    g.n16[c] = k.n16[c];
    g.nz16[c] = arena + at;
    memcpy(arena + at, k.nz16[c], k.n16[c] * sizeof(uint32_t));
    at += k.n16[c];
  }
  at += at & 1;
  for (int c = 0; c < 3; c++) {
    g.n32[c] = k.n32[c];
    g.nz32[c] = arena + at;
    memcpy(arena + at, k.nz32[c], 2 * static_cast<size_t>(k.n32[c]) * sizeof(uint32_t));
    at += 2 * static_cast<size_t>(k.n32[c]);
  }
  if (jxlgpu_submit_groups_sparse(be.ctx, 1, &g, 0) != JXLGPU_OK)
    return JXL_FAILURE("gpu backend: submit_groups_sparse: %s", jxlgpu_last_error(be.ctx));
  return true;
}

inline jxl::Status GroupDecoded(jxl::PassesDecoderState* ds, size_t group) {
  GpuBackend& be = GpuBackend::Get();
  GpuFrame* fr = be.Find(ds);
  if (!fr || !fr->begun) return JXL_FAILURE("gpu backend: group before frame_begin");
  if (fr->sparse) return GroupDecodedSparse(be, fr, ds, group);
  const uint32_t row = static_cast<uint32_t>(group / fr->xg);
  if (fr->row_count[row].fetch_add(1) + 1 != fr->xg) return true;
  // this thread completed the row: hand the whole row over (adjacent [3][65536] blocks -> one DMA)
  const jxl::FrameDimensions& d = ds->shared->frame_dim;
  std::vector<uint32_t> idx(fr->xg);
  std::vector<const void*> co(3 * fr->xg);
  std::vector<size_t> nco(fr->xg);
  for (uint32_t gx = 0; gx < fr->xg; gx++) {
    const size_t g = static_cast<size_t>(row) * fr->xg + gx;
    idx[gx] = static_cast<uint32_t>(g);
    const jxl::Rect br = d.BlockGroupRect(g);
    nco[gx] = 64 * br.xsize() * br.ysize();
    for (size_t c = 0; c < 3; c++) {
      jxl::ACPtr p = ds->coefficients->PlaneRow(c, g, 0);
      co[3 * gx + c] = fr->is16 ? static_cast<const void*>(p.ptr16) : static_cast<const void*>(p.ptr32);
    }
  }
  if (jxlgpu_submit_groups(be.ctx, fr->xg, idx.data(), 0, co.data(), nco.data()) != JXLGPU_OK) {
    fr->error.store(1);
    return JXL_FAILURE("gpu backend: submit_groups: %s", jxlgpu_last_error(be.ctx));
  }
  return true;
}

// ---- site 4: FinalizeFrame ----
inline jxl::Status FinishFrame(jxl::PassesDecoderState* ds) {
  GpuBackend& be = GpuBackend::Get();
  GpuFrame* fr = be.Find(ds);
  if (!fr) return true;
  jxl::Status st = true;
  if (fr->begun) {
    if (jxlgpu_frame_finish(be.ctx, ds->main_output.buffer, ds->main_output.stride) != JXLGPU_OK)
      st = JXL_FAILURE("gpu backend: frame_finish: %s", jxlgpu_last_error(be.ctx));
    else be.frames_taken++;
  }
  std::lock_guard<std::mutex> lk(be.mu);
  be.frames.erase(ds);
  return st;
}

bool DontDraw(const jxl::PassesDecoderState* dec_state) { return FrameActive(dec_state); }

#endif  // JXLB_GPU_BACKEND_IMPLEMENTATION

}  // namespace jxlb_integration

// exported by the patched library so that a test / bench can see that frames really took the GPU path
extern "C" __attribute__((visibility("default"))) unsigned long long jxlb_gpu_backend_frames_taken(void);

#endif  // JXL_B200_INTEGRATION_LIBJXL_GPU_BACKEND_H_
