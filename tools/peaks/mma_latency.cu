// Dependent-issue latency and single-warp issue interval of mma.sync.m16n8k8.tf32 on sm_100a: one warp per SM
// sub-partition (block = 128, one CTA per SM) runs NACC independent accumulator chains; cycles per HMMA by clock64.
// NACC = 1 gives the dependent (accumulate-into-itself) latency; large NACC the single-warp issue interval.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NACC>
__global__ void __launch_bounds__(128) k(float* out, long long* cyc, int iters) {
  float c[NACC][4];
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) mma_tf32(c[i], a, b);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678f) out[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NACC> void run(float* out, long long* cyc, int warps) {
  const int iters = 2000;
  k<NACC><<<148, 32 * warps>>>(out, cyc, iters);
  k<NACC><<<148, 32 * warps>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"bench\": \"hmma_tf32_chain\", \"nacc\": %d, \"warps_per_sm\": %d, \"cycles_per_hmma_per_warp\": %.2f}\n", NACC, warps, (double)h / (iters * 8.0 * NACC));
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  for (int warps : {4}) {  // the kernel is bounded to 128 threads: one warp per sub-partition run<1>(out, cyc, warps); run<2>(out, cyc, warps); run<3>(out, cyc, warps); run<4>(out, cyc, warps); run<6>(out, cyc, warps); run<8>(out, cyc, warps); run<12>(out, cyc, warps); }
  return 0;
}
