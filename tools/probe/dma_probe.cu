// DMA behaviour probe: many medium H2D copies (one per AC group) with / without a concurrent D2H stream.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t chunk = 3 * 65536 * 4, n = 510, big = 23 << 20;
  char *h, *ho, *d, *dd;
  cudaHostAlloc(&h, chunk * n, 0); cudaHostAlloc(&ho, big * 17, 0);
  cudaMalloc(&d, chunk * n); cudaMalloc(&dd, big * 17);
  cudaStream_t s[8], sd; for (auto& x : s) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking);
  for (int ns2 : {1, 2, 4}) {
    for (int rep = 0; rep < 3; rep++) {
      cudaDeviceSynchronize();
      double t0 = now();
      for (size_t i = 0; i < n; i++) cudaMemcpyAsync(d + i * chunk, h + i * chunk, chunk, cudaMemcpyHostToDevice, s[i % ns2]);
      for (int i = 0; i < 17; i++) cudaMemcpyAsync(ho + i * big, dd + i * big, big, cudaMemcpyDeviceToHost, sd);
      cudaDeviceSynchronize();
      if (rep == 2) printf("H2D 510 chunks on %d stream(s) + D2H 17x23MB: %.2f ms\n", ns2, (now() - t0) * 1e3);
    }
  }
  for (int rep = 0; rep < 3; rep++) {
    cudaDeviceSynchronize();
    double t0 = now();
    for (size_t i = 0; i < 17; i++) cudaMemcpyAsync(d + i * 30 * chunk, h + i * 30 * chunk, 30 * chunk, cudaMemcpyHostToDevice, s[0]);
    for (int i = 0; i < 17; i++) cudaMemcpyAsync(ho + i * big, dd + i * big, big, cudaMemcpyDeviceToHost, sd);
    cudaDeviceSynchronize();
    if (rep == 2) printf("H2D 17 row-sized copies (1 stream) + D2H 17x23MB: %.2f ms\n", (now() - t0) * 1e3);
  }
  for (int rep = 0; rep < 3; rep++) {
    cudaDeviceSynchronize();
    double t0 = now();
    for (size_t i = 0; i < n; i++) cudaMemcpyAsync(d + i * chunk, h + i * chunk, chunk, cudaMemcpyHostToDevice, s[0]);
    for (int i = 0; i < 17 * 16; i++) cudaMemcpyAsync(ho + i * (big / 16), dd + i * (big / 16), big / 16, cudaMemcpyDeviceToHost, sd);
    cudaDeviceSynchronize();
    if (rep == 2) printf("H2D 510 chunks (1 stream) + D2H 272 x 1.4MB: %.2f ms\n", (now() - t0) * 1e3);
  }
  for (int mode = 0; mode < 0; mode++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaDeviceSynchronize();
      double t0 = now();
      int ns = (mode == 0) ? 1 : 8;
      if (mode == 4) { cudaMemcpyAsync(d, h, chunk * n, cudaMemcpyHostToDevice, s[0]); }
      else if (mode != 3) for (size_t i = 0; i < n; i++) cudaMemcpyAsync(d + i * chunk, h + i * chunk, chunk, cudaMemcpyHostToDevice, s[i % ns]);
      if (mode >= 2 && mode != 4) for (int i = 0; i < 17; i++) cudaMemcpyAsync(ho + i * big, dd + i * big, big, cudaMemcpyDeviceToHost, sd);
      cudaDeviceSynchronize();
      double t = now() - t0;
      if (rep == 2) printf("mode %d (%s): %.2f ms\n", mode,
          mode == 0 ? "510 x 786KB H2D, 1 stream" : mode == 1 ? "510 x 786KB H2D, 8 streams" :
          mode == 2 ? "H2D 8 streams + 17 x 23MB D2H concurrently" : mode == 3 ? "17 x 23MB D2H only" : "one 401MB H2D", t * 1e3);
    }
  }
  return 0;
}
