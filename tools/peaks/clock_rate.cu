// SM clock actually delivered: clock64 ticks per globaltimer nanosecond, for a small (32 CTAs x 128 threads) and a
// full (148 x 4 CTAs) launch, short and long.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, long long spin) {
  long long c0 = clock64(); unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  float x = threadIdx.x;
  while (clock64() - c0 < spin) x = x * 1.0001f + 0.5f;
  long long c1 = clock64(); unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = c1 - c0; out[1] = (long long)(g1 - g0); out[2] = (long long)x; }
}
int main() {
  long long* d; cudaMalloc(&d, 64); long long h[3];
  for (int rep = 0; rep < 3; ++rep)
    for (int grid : {32, 592}) for (long long spin : {100000ll, 2000000ll, 200000000ll}) {
      k<<<grid, 128>>>(d, spin); cudaDeviceSynchronize(); cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
      printf("{\"bench\": \"clock_rate\", \"grid\": %d, \"spin_cycles\": %lld, \"sm_mhz\": %.1f}\n", grid, spin, 1e3 * h[0] / (double)h[1]);
    }
  return 0;
}
