// Pinned device->host bandwidth of strided (cudaMemcpy2DAsync) copies of the live column ranges of obs[N][D] against one
// contiguous copy of the whole array: can the always-zero observation columns (contact columns of the frozen bodies)
// stay off the PCIe link without losing DMA efficiency?
// nvcc -O3 -o d2h_2d_probe d2h_2d_probe.cu && ./d2h_2d_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <initializer_list>
int main() {
  const size_t N = 1 << 20, D = 114, pitch = D * 4;
  char *dev, *host;
  cudaMalloc(&dev, N * pitch); cudaMemset(dev, 1, N * pitch);
  cudaMallocHost(&host, N * pitch);
  cudaStream_t s[4]; for (auto& x : s) cudaStreamCreate(&x);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto fn, double bytes) {
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, s[0]);
      fn();
      for (int i = 1; i < 4; ++i) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); cudaEventRecord(e, s[i]); cudaStreamWaitEvent(s[0], e, 0); cudaEventDestroy(e); }
      cudaEventRecord(e1, s[0]); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best;
    }
    printf("%-58s %8.3f ms  %6.1f GB/s payload  (%.0f MB)\n", name, best, bytes / best / 1e6, bytes / 1e6);
  };
  timeit("contiguous N x 456 B", [&] { cudaMemcpyAsync(host, dev, N * pitch, cudaMemcpyDeviceToHost, s[0]); }, (double)N * pitch);
  for (size_t w : {448u, 396u, 336u, 256u, 224u, 172u, 128u, 108u, 64u, 32u, 4u}) {
    char name[96]; snprintf(name, sizeof name, "2D width %zu B of pitch 456 B, one stream", w);
    timeit(name, [&] { cudaMemcpy2DAsync(host, pitch, dev, pitch, w, N, cudaMemcpyDeviceToHost, s[0]); }, (double)N * w);
  }
  timeit("2D 224 B + 2D 108 B @284 + 2D 4 B @452, one stream", [&] {
    cudaMemcpy2DAsync(host, pitch, dev, pitch, 224, N, cudaMemcpyDeviceToHost, s[0]);
    cudaMemcpy2DAsync(host + 284, pitch, dev + 284, pitch, 108, N, cudaMemcpyDeviceToHost, s[0]);
    cudaMemcpy2DAsync(host + 452, pitch, dev + 452, pitch, 4, N, cudaMemcpyDeviceToHost, s[0]); }, (double)N * 336);
  timeit("2D 224 B + 2D 172 B @284, one stream", [&] {
    cudaMemcpy2DAsync(host, pitch, dev, pitch, 224, N, cudaMemcpyDeviceToHost, s[0]);
    cudaMemcpy2DAsync(host + 284, pitch, dev + 284, pitch, 172, N, cudaMemcpyDeviceToHost, s[0]); }, (double)N * 396);
  timeit("2D 224 B + 2D 172 B @284, two streams", [&] {
    cudaMemcpy2DAsync(host, pitch, dev, pitch, 224, N, cudaMemcpyDeviceToHost, s[0]);
    cudaMemcpy2DAsync(host + 284, pitch, dev + 284, pitch, 172, N, cudaMemcpyDeviceToHost, s[1]); }, (double)N * 396);
  // compact device buffer [N][84] -> contiguous copy (the consumer would need a host-side expansion)
  timeit("contiguous N x 336 B (compact layout)", [&] { cudaMemcpyAsync(host, dev, N * 336, cudaMemcpyDeviceToHost, s[0]); }, (double)N * 336);
  // compact device rows scattered into the full-pitch host array by the DMA engine
  timeit("2D: compact dev rows (pitch 224) -> host pitch 456", [&] { cudaMemcpy2DAsync(host, pitch, dev, 224, 224, N, cudaMemcpyDeviceToHost, s[0]); }, (double)N * 224);
  return 0;
}
