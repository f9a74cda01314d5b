// tests/emu/cuda_fp16.h -- TEST INFRASTRUCTURE (see cuda_runtime.h here): binary16 storage type and the
// one conversion the kernels use, IEEE round-to-nearest-even.
#pragma once
#include <cstdint>
#include <cstring>
struct __half { uint16_t v; };
inline __half __float2half_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u, a = u & 0x7fffffffu;
  __half h;
  if (a >= 0x7f800000u) { h.v = (uint16_t)(sign | 0x7c00u | (a > 0x7f800000u ? 0x200u | ((a >> 13) & 0x3ffu) : 0u)); return h; }
  if (a >= 0x477ff000u) { h.v = (uint16_t)(sign | 0x7c00u); return h; }
  if (a < 0x33000001u) { h.v = (uint16_t)sign; return h; }
  const int e = (int)(a >> 23) - 127;
  const uint32_t m = (a & 0x7fffffu) | 0x800000u;
  int shift = 13;
  uint32_t he = (uint32_t)(e + 15);
  if (e < -14) { shift += -14 - e; he = 0; }
  const uint32_t halfway = 1u << (shift - 1), rem = m & ((1u << shift) - 1u);
  uint32_t r = m >> shift;
  if (rem > halfway || (rem == halfway && (r & 1u))) r++;
  h.v = (uint16_t)(sign | (he ? ((he - 1u) << 10) + r : r));
  return h;
}
