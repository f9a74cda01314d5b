// host_feed.cc -- the C ABI of include/jxl_b200.h driven from plain C++ the way libjxl's FrameDecoder
// would drive it: one frame_begin on the coordinating thread, worker threads that
// hand over entropy-decoded AC groups as they finish (here: read from a dump file, in shuffled order),
// frame_finish on the coordinating thread.  No CUDA headers, no Python: this is all a host needs.
//
//   g++ -O2 -std=c++17 -Iinclude examples/host_feed.cc -Llibjxl_b200 -ljxl_b200 -pthread -o host_feed
//   LD_LIBRARY_PATH=libjxl_b200 ./host_feed frame.bin out.raw [threads] [dense|sparse]
//
// frame.bin is written by examples/dump_frame.py (a jxlgpu_frame with its planes, then the coefficient
// groups); out.raw receives the pixels in the frame's out_format.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "jxl_b200.h"

namespace {

struct Reader {
  FILE* f;
  template <typename T>
  std::vector<T> array() {
    uint64_t n = 0;
    if (fread(&n, 8, 1, f) != 1) die("truncated dump");
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) die("truncated dump");
    return v;
  }
  [[noreturn]] static void die(const char* what) {
    fprintf(stderr, "host_feed: %s\n", what);
    exit(2);
  }
};

size_t out_pixel_bytes(uint32_t fmt) {
  switch (fmt) {
    case JXLGPU_OUT_RGB_F32: return 12;
    case JXLGPU_OUT_PLANAR_F32: return 4;
    case JXLGPU_OUT_RGB_U8: return 3;
    case JXLGPU_OUT_RGBA_U8: return 4;
    default: return 6;
  }
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: host_feed frame.bin out.raw [threads] [dense|sparse]\n");
    return 2;
  }
  const unsigned threads = argc > 3 ? std::max(1, atoi(argv[3])) : 4;
  const bool sparse = argc > 4 && std::string(argv[4]) == "sparse";
  Reader r{fopen(argv[1], "rb")};
  if (!r.f) Reader::die("cannot open the dump");

  // ---- the frame description, as examples/dump_frame.py wrote it ----
  jxlgpu_frame fr;
  if (fread(&fr, sizeof(fr), 1, r.f) != 1) Reader::die("truncated dump");
  auto acs = r.array<uint8_t>();
  auto quant = r.array<int32_t>();
  auto sharp = r.array<uint8_t>();
  auto ytox = r.array<int8_t>();
  auto ytob = r.array<int8_t>();
  auto dc = r.array<float>();
  auto dq = r.array<float>();
  const size_t xb = fr.xsize_blocks, yb = fr.ysize_blocks;
  fr.ac_strategy = acs.data();
  fr.raw_quant = quant.data();
  fr.epf_sharpness = sharp.empty() ? nullptr : sharp.data();
  fr.ytox_map = ytox.data();
  fr.ytob_map = ytob.data();
  for (int c = 0; c < 3; c++) fr.dc[c] = dc.data() + c * xb * yb;
  fr.dequant_table = dq.data();
  fr.quant_dc[0] = fr.quant_dc[1] = fr.quant_dc[2] = nullptr;
  fr.dc_group_mul = nullptr;
  if (fr.upsampling > 1) Reader::die("this example does not carry the upsampling weights of the dump");
  fr.upsampling_weights = nullptr;

  const size_t es = fr.ac_type == JXLGPU_AC_INT16 ? 2 : 4;
  const uint32_t xg = (fr.xsize + 255) / 256, yg = (fr.ysize + 255) / 256, num_groups = xg * yg;

  if (getenv("HOST_FEED_PARSE_ONLY")) {  // (tests: check the reader without a device)
    uint64_t nonzero = 0, bytes = 0;
    for (uint32_t g = 0; g < num_groups; g++) {
      const size_t nbx = std::min<size_t>(32, xb - (g % xg) * 32), nby = std::min<size_t>(32, yb - (g / xg) * 32);
      std::vector<uint8_t> buf(64 * nbx * nby * es);
      for (int c = 0; c < 3; c++) {
        if (fread(buf.data(), 1, buf.size(), r.f) != buf.size()) Reader::die("truncated coefficients");
        bytes += buf.size();
        for (size_t k = 0; k < buf.size() / es; k++)
          nonzero += (es == 2 ? ((const int16_t*)buf.data())[k] : ((const int32_t*)buf.data())[k]) != 0;
      }
    }
    printf("parsed %ux%u blocks=%zux%zu groups=%u coeff_bytes=%llu nonzero=%llu dc0=%.9g dq_last=%.9g\n", fr.xsize, fr.ysize,
           xb, yb, num_groups, (unsigned long long)bytes, (unsigned long long)nonzero, (double)dc[0], (double)dq.back());
    return fgetc(r.f) == EOF ? 0 : 7;
  }

  jxlgpu_config cfg = {JXLGPU_ABI_VERSION, 0, threads, 0};
  jxlgpu_ctx* ctx = nullptr;
  int rc = jxlgpu_create(&ctx, &cfg);
  if (rc != JXLGPU_OK) {  // a real host would now take its CPU path; there is none inside the library
    fprintf(stderr, "host_feed: jxlgpu_create: %s\n", jxlgpu_error_string(rc));
    return 3;
  }

  // ---- coefficient groups in page-locked memory: [group][channel][65536] ----
  void* pinned = jxlgpu_alloc_pinned((size_t)num_groups * 3 * JXLGPU_GROUP_COEFFS * es);
  if (!pinned) Reader::die("jxlgpu_alloc_pinned failed");
  std::vector<size_t> ncoeff(num_groups);
  for (uint32_t g = 0; g < num_groups; g++) {
    const size_t nbx = std::min<size_t>(32, xb - (g % xg) * 32), nby = std::min<size_t>(32, yb - (g / xg) * 32);
    ncoeff[g] = 64 * nbx * nby;
    for (int c = 0; c < 3; c++) {
      uint8_t* dst = (uint8_t*)pinned + ((size_t)g * 3 + c) * JXLGPU_GROUP_COEFFS * es;
      if (fread(dst, es, ncoeff[g], r.f) != ncoeff[g]) Reader::die("truncated coefficients");
    }
  }
  fclose(r.f);

  // sparse hand-off: what an entropy decoder that appends non-zeros instead of scattering would have
  // produced; built here from the dense planes (outside anything one would time)
  std::vector<std::vector<uint32_t>> nz16(num_groups * 3), nz32(num_groups * 3);
  if (sparse) {
    for (uint32_t g = 0; g < num_groups; g++)
      for (int c = 0; c < 3; c++) {
        const uint8_t* src = (const uint8_t*)pinned + ((size_t)g * 3 + c) * JXLGPU_GROUP_COEFFS * es;
        for (size_t k = 0; k < ncoeff[g]; k++) {
          const int32_t v = es == 2 ? ((const int16_t*)src)[k] : ((const int32_t*)src)[k];
          if (!v) continue;
          if (v >= -32768 && v <= 32767) {
            nz16[g * 3 + c].push_back((uint32_t)k << 16 | (uint16_t)v);
          } else {
            nz32[g * 3 + c].push_back((uint32_t)k);
            nz32[g * 3 + c].push_back((uint32_t)v);
          }
        }
      }
  }

  const size_t rows = fr.out_format == JXLGPU_OUT_PLANAR_F32 ? 3 * (size_t)fr.ysize : fr.ysize;
  const size_t stride = (size_t)fr.xsize * out_pixel_bytes(fr.out_format);
  void* out = jxlgpu_alloc_pinned(rows * stride);
  if (!out) Reader::die("jxlgpu_alloc_pinned failed");

  // ---- one frame ----
  if ((rc = jxlgpu_frame_begin(ctx, &fr)) != JXLGPU_OK || (rc = jxlgpu_frame_set_output(ctx, out, stride)) != JXLGPU_OK) {
    fprintf(stderr, "host_feed: frame_begin: %s (%s)\n", jxlgpu_error_string(rc), jxlgpu_last_error(ctx));
    return 4;
  }
  std::vector<uint32_t> order(num_groups);
  for (uint32_t g = 0; g < num_groups; g++) order[g] = g;
  std::shuffle(order.begin(), order.end(), std::mt19937(1234));  // groups finish in any order
  std::atomic<uint32_t> next{0};
  std::atomic<int> failed{0};
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < threads; t++)
    pool.emplace_back([&, t] {
      for (uint32_t i; (i = next.fetch_add(1)) < num_groups;) {
        const uint32_t g = order[i];
        int e;
        if (sparse) {
          jxlgpu_sparse_group sg = {};
          sg.group_idx = g;
          for (int c = 0; c < 3; c++) {
            sg.n16[c] = (uint32_t)nz16[g * 3 + c].size();
            sg.n32[c] = (uint32_t)(nz32[g * 3 + c].size() / 2);
            sg.nz16[c] = nz16[g * 3 + c].data();
            sg.nz32[c] = nz32[g * 3 + c].data();
          }
          e = jxlgpu_submit_groups_sparse(ctx, 1, &sg, t);
        } else {
          const void* co[3];
          for (int c = 0; c < 3; c++) co[c] = (const uint8_t*)pinned + ((size_t)g * 3 + c) * JXLGPU_GROUP_COEFFS * es;
          e = jxlgpu_submit_group(ctx, g, t, co, ncoeff[g]);
        }
        if (e != JXLGPU_OK) failed = e;
      }
    });
  for (auto& th : pool) th.join();
  if (failed) {
    fprintf(stderr, "host_feed: submit: %s (%s)\n", jxlgpu_error_string(failed), jxlgpu_last_error(ctx));
    return 5;
  }
  if ((rc = jxlgpu_frame_finish(ctx, out, stride)) != JXLGPU_OK) {
    fprintf(stderr, "host_feed: frame_finish: %s (%s)\n", jxlgpu_error_string(rc), jxlgpu_last_error(ctx));
    return 6;
  }
  FILE* fo = fopen(argv[2], "wb");
  if (!fo || fwrite(out, stride, rows, fo) != rows) Reader::die("cannot write the output");
  fclose(fo);
  printf("host_feed: %ux%u, %u groups, %u threads, %s hand-off, %llu kernel launches\n", fr.xsize, fr.ysize, num_groups,
         threads, sparse ? "sparse" : "dense", (unsigned long long)jxlgpu_launch_count(ctx));
  jxlgpu_free_pinned(out);
  jxlgpu_free_pinned(pinned);
  jxlgpu_destroy(ctx);
  return 0;
}
