// Dependent-load latency vs footprint on the device (pointer chase, one thread), and the same chase with 1 warp x 32
// independent chains.  nvcc -O3 -arch=sm_100a -o chase chase.cu && ./chase
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
__global__ void chase(const unsigned* __restrict__ next, unsigned start, int hops, unsigned* out, long long* cyc) {
  unsigned p = start;
  long long t0 = clock64();
  for (int i = 0; i < hops; ++i) p = next[p];
  long long t1 = clock64();
  *out = p; *cyc = t1 - t0;
}
int main() {
  const size_t sizes_mb[] = {1, 32, 100, 512, 2048, 8192, 32768};
  for (size_t mb : sizes_mb) {
    const size_t n = mb * 1024 * 1024 / 4;
    const size_t stride = 64 * 1024 / 4;  // one hop per 64 KB region -> a new 2 MB page every 32 hops on average (random)
    const size_t slots = n / stride;
    std::vector<unsigned> order(slots);
    for (size_t i = 0; i < slots; ++i) order[i] = (unsigned)i;
    std::mt19937 rng(1);
    std::shuffle(order.begin(), order.end(), rng);
    unsigned* d; cudaMalloc(&d, n * 4); cudaMemset(d, 0, n * 4);
    std::vector<unsigned> host(slots);
    // next[order[i]*stride] = order[i+1]*stride
    for (size_t i = 0; i < slots; ++i) {
      unsigned v = (unsigned)(order[(i + 1) % slots] * stride);
      cudaMemcpy(d + order[i] * stride, &v, 4, cudaMemcpyHostToDevice);
    }
    unsigned* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    const int hops = (int)std::min<size_t>(slots, 4096);
    for (int rep = 0; rep < 3; ++rep) {
      chase<<<1, 1>>>(d, (unsigned)(order[0] * stride), hops, out, cyc);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      if (rep == 2) printf("footprint %6zu MB: %d hops, %.0f cycles/hop (~%.0f ns at 1.9 GHz)\n", mb, hops, (double)c / hops, (double)c / hops / 1.9);
    }
    cudaFree(d); cudaFree(out); cudaFree(cyc);
  }
  return 0;
}
