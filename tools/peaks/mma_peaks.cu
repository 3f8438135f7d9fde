// Legacy warp-level mma.sync throughput on sm_100a (B200): is the 3xTF32 / split-bf16 route for the small
// MLP layers faster than the FP32 FMA pipe?  (tcgen05 is the native tensor path; mma.sync is what the
// north star names for the warp-level variant.)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void dmma_884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma_16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int KIND, int NACC>
__global__ void __launch_bounds__(256) k_dmma(float* out, int iters) {
  double c[NACC][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0; }
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 0.5 + threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (KIND == 0) { double cc[2] = {c[i][0], c[i][1]}; dmma_884(cc, a[0], b[0]); c[i][0] = cc[0]; c[i][1] = cc[1]; }
        else if (KIND == 1) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; dmma_1688(c[i], aa, bb); }
        else dmma_16816(c[i], a, b);
      }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = (float)s;
}

template <int KIND, int NACC>
__global__ void __launch_bounds__(256) k_mma(float* out, int iters) {
  float c[NACC][4];
  uint32_t a[4], b[2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
  a[0] = 0x3f800000u + threadIdx.x; a[1] = 0x3f000000u; a[2] = 0x3e800000u; a[3] = 0x3f800000u;
  b[0] = 0x3f800000u; b[1] = 0x3f000000u + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (KIND == 0) mma_tf32(c[i], a, b);
        else if (KIND == 1) mma_bf16(c[i], a, b);
        else mma_f16(c[i], a, b);
      }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678f) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  float* out; CK(cudaMalloc(&out, 4));
  const int iters = 2000;
  for (int bps : {1, 2, 4}) {
    int grid = sms * bps;
#define RUN(KIND, NACC, NAME, MACS) { double ms = time_ms([&] { k_mma<KIND, NACC><<<grid, 256>>>(out, iters); }); \
      double fl = 2.0 * MACS * (double)grid * 8 * iters * 4 * NACC; \
      printf("{\"bench\": \"%s\", \"nacc\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.1f}\n", NAME, NACC, bps, ms, fl / ms * 1e-9); }
    RUN(0, 4, "mma_sync_m16n8k8_tf32", 1024.0) RUN(0, 8, "mma_sync_m16n8k8_tf32", 1024.0)
    RUN(1, 4, "mma_sync_m16n8k16_bf16", 2048.0) RUN(1, 8, "mma_sync_m16n8k16_bf16", 2048.0)
    RUN(2, 8, "mma_sync_m16n8k16_f16", 2048.0)
#define RUND(KIND, NACC, NAME, MACS) { double ms = time_ms([&] { k_dmma<KIND, NACC><<<grid, 256>>>(out, iters / 4); }); \
      double fl = 2.0 * MACS * (double)grid * 8 * (iters / 4) * 4 * NACC; \
      printf("{\"bench\": \"%s\", \"nacc\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.1f}\n", NAME, NACC, bps, ms, fl / ms * 1e-9); }
    RUND(0, 8, "dmma_m8n8k4_f64", 256.0) RUND(1, 8, "dmma_m16n8k8_f64", 1024.0) RUND(2, 8, "dmma_m16n8k16_f64", 2048.0)
  }
  return 0;
}
