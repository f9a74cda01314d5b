// Is rcp.approx + one FMA Newton step equal to the correctly rounded reciprocal for every
// integer-valued float the dequantiser can see (|q| in [2, 2^24], and larger even integers)?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float rcp_nr(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  const float e = fmaf(-x, r, 1.0f);
  return fmaf(r, e, r);
}
__global__ void k(unsigned long long* bad, unsigned* first) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 2^24 + extra
  float x;
  if (i < (1u << 24)) x = (float)(i + 2);
  else x = __int_as_float(0x4e800000 + (i - (1u << 24)) * 8);  // exponent 2^30, every 8th significand
  for (int sgn = 0; sgn < 2; sgn++) {
    const float v = sgn ? -x : x;
    if (rcp_nr(v) != __frcp_rn(v)) {
      atomicAdd(bad, 1ull);
      atomicMin(first, i);
    }
  }
}
int main() {
  unsigned long long* bad; unsigned* first;
  cudaMallocManaged(&bad, 8); cudaMallocManaged(&first, 4);
  *bad = 0; *first = 0xffffffffu;
  k<<<(1u << 24) / 256 + 4096, 256>>>(bad, first);
  cudaDeviceSynchronize();
  printf("mismatches: %llu first index %u\n", *bad, *first);
  return 0;
}
