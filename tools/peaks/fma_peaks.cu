// FP32 / FP32x2 / FP64 FMA-pipe peak micro-benchmarks for sm_100a (B200).
//
// MEASURED_PEAKS.json (driver-written) has HBM and bf16 tensor peaks only; the MPPI rollout is
// bound by the FP32 (or FP64) FMA pipe, so the roofline denominator for it has to be measured
// here.  Also measures the operand-delivery variants the rollout kernel can choose from:
//   * weights from registers (upper bound), * weights as constant-bank operands,
//   * weights by broadcast LDS.128 from shared memory at several FMA:LDS ratios.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fma_peaks fma_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}

constexpr int ILP = 16;

__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b) {
  uint64_t acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = pack2(threadIdx.x * 1e-3f + i, i);
  uint64_t a2 = pack2(a, a), b2 = pack2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = ffma2(acc[i], a2, b2);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float x, y; unpack2(acc[i], x, y); s += x + y; }
  if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dfma(float* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = (float)s;
}

// ---- constant-bank operand: acc[j] += x * c[i*ILP + j], cycling through NCONST weights --------
constexpr int NCONST = 4096;  // 16 KB, about the push REX model (3695 params)
__constant__ float c_w[NCONST];

__global__ void __launch_bounds__(256) k_ffma_const(float* out, int iters, float x0) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  float x = x0 + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCONST / ILP; ++i) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) acc[j] = fmaf(x, c_w[i * ILP + j], acc[j]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

// ---- shared-memory broadcast weights -------------------------------------------------------
// S samples per thread; each LDS.128 fetches 4 weights (outputs j..j+3 for input i) for all lanes.
// scalar: 4*S FFMA per LDS.128;  packed: 2*S FFMA2 per LDS.128 (x duplicated into a pair).
template <int S>
__global__ void __launch_bounds__(128) k_lds_ffma(float* out, int iters, float x0) {
  __shared__ float4 w[NCONST / 4];
  for (int i = threadIdx.x; i < NCONST / 4; i += blockDim.x) w[i] = make_float4(1e-3f * i, 2e-3f, 3e-3f, 4e-3f);
  __syncthreads();
  float acc[S][32];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[s][j] = j + s;
  float x[S];
#pragma unroll
  for (int s = 0; s < S; ++s) x[s] = x0 + threadIdx.x * 1e-6f + s;
  for (int it = 0; it < iters; ++it) {
    // one "layer" = 32 outputs x (NCONST/32) inputs
#pragma unroll 4
    for (int i = 0; i < NCONST / 32; ++i) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        float4 ww = w[i * 8 + j4];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          acc[s][j4 * 4 + 0] = fmaf(x[s], ww.x, acc[s][j4 * 4 + 0]);
          acc[s][j4 * 4 + 1] = fmaf(x[s], ww.y, acc[s][j4 * 4 + 1]);
          acc[s][j4 * 4 + 2] = fmaf(x[s], ww.z, acc[s][j4 * 4 + 2]);
          acc[s][j4 * 4 + 3] = fmaf(x[s], ww.w, acc[s][j4 * 4 + 3]);
        }
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < 32; ++j) r += acc[s][j];
  if (r == 12345.678f) out[0] = r;
}

template <int S>
__global__ void __launch_bounds__(128) k_lds_ffma2(float* out, int iters, float x0) {
  __shared__ ulonglong2 w[NCONST / 4];
  for (int i = threadIdx.x; i < NCONST / 4; i += blockDim.x) {
    ulonglong2 v; v.x = pack2(1e-3f * i, 2e-3f); v.y = pack2(3e-3f, 4e-3f); w[i] = v;
  }
  __syncthreads();
  uint64_t acc[S][16];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[s][j] = pack2(j, s);
  uint64_t x[S];
#pragma unroll
  for (int s = 0; s < S; ++s) { float v = x0 + threadIdx.x * 1e-6f + s; x[s] = pack2(v, v); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int i = 0; i < NCONST / 32; ++i) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        ulonglong2 ww = w[i * 8 + j4];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          acc[s][j4 * 2 + 0] = ffma2(x[s], ww.x, acc[s][j4 * 2 + 0]);
          acc[s][j4 * 2 + 1] = ffma2(x[s], ww.y, acc[s][j4 * 2 + 1]);
        }
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < 16; ++j) { float a, b; unpack2(acc[s][j], a, b); r += a + b; }
  if (r == 12345.678f) out[0] = r;
}

// FP64 with broadcast LDS.128 (2 weights per load), S samples per thread
template <int S>
__global__ void __launch_bounds__(128) k_lds_dfma(float* out, int iters, double x0) {
  __shared__ double2 w[NCONST / 2];
  for (int i = threadIdx.x; i < NCONST / 2; i += blockDim.x) w[i] = make_double2(1e-3 * i, 2e-3);
  __syncthreads();
  double acc[S][32];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[s][j] = j + s;
  double x[S];
#pragma unroll
  for (int s = 0; s < S; ++s) x[s] = x0 + threadIdx.x * 1e-6 + s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 2
    for (int i = 0; i < NCONST / 32; ++i) {
#pragma unroll
      for (int j2 = 0; j2 < 16; ++j2) {
        double2 ww = w[i * 16 + j2];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          acc[s][j2 * 2 + 0] = fma(x[s], ww.x, acc[s][j2 * 2 + 0]);
          acc[s][j2 * 2 + 1] = fma(x[s], ww.y, acc[s][j2 * 2 + 1]);
        }
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < 32; ++j) r += acc[s][j];
  if (r == 12345.678) out[0] = (float)r;
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
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  float* out; CK(cudaMalloc(&out, 4));
  std::vector<float> hw(NCONST); for (int i = 0; i < NCONST; ++i) hw[i] = 1e-3f * (i % 7);
  CK(cudaMemcpyToSymbol(c_w, hw.data(), NCONST * 4));
  const int bps[] = {1, 2, 4, 8};
  for (int bi = 0; bi < 4; ++bi) {
    int grid = sms * bps[bi];
    {
      int iters = 4000; double ms = time_ms([&] { k_ffma<<<grid, 256>>>(out, iters, 1.0001f, 1e-7f); });
      double fl = 2.0 * grid * 256.0 * iters * 8 * ILP;
      printf("{\"bench\": \"ffma_reg\", \"blocks_per_sm\": %d, \"threads\": 256, \"ms\": %.4f, \"tflops\": %.2f}\n", bps[bi], ms, fl / ms * 1e-9);
    }
    {
      int iters = 4000; double ms = time_ms([&] { k_ffma2<<<grid, 256>>>(out, iters, 1.0001f, 1e-7f); });
      double fl = 4.0 * grid * 256.0 * iters * 8 * ILP;
      printf("{\"bench\": \"ffma2_reg\", \"blocks_per_sm\": %d, \"threads\": 256, \"ms\": %.4f, \"tflops\": %.2f}\n", bps[bi], ms, fl / ms * 1e-9);
    }
    {
      int iters = 2000; double ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0001, 1e-7); });
      double fl = 2.0 * grid * 256.0 * iters * 8 * ILP;
      printf("{\"bench\": \"dfma_reg\", \"blocks_per_sm\": %d, \"threads\": 256, \"ms\": %.4f, \"tflops\": %.2f}\n", bps[bi], ms, fl / ms * 1e-9);
    }
    {
      int iters = 200; double ms = time_ms([&] { k_ffma_const<<<grid, 256>>>(out, iters, 1e-3f); });
      double fl = 2.0 * grid * 256.0 * iters * NCONST;
      printf("{\"bench\": \"ffma_const16k\", \"blocks_per_sm\": %d, \"threads\": 256, \"ms\": %.4f, \"tflops\": %.2f}\n", bps[bi], ms, fl / ms * 1e-9);
    }
  }
  const int bps2[] = {1, 2, 4, 6, 8};
  for (int bi = 0; bi < 5; ++bi) {
    int grid = sms * bps2[bi];
    int iters = 200;
#define RUN_LDS(NAME, KERN, S, FPF)                                                           \
    {                                                                                         \
      double ms = time_ms([&] { KERN<S><<<grid, 128>>>(out, iters, 1e-3f); });                \
      double fl = 2.0 * grid * 128.0 * iters * (double)NCONST * S;                            \
      printf("{\"bench\": \"%s\", \"samples_per_thread\": %d, \"blocks_per_sm\": %d, \"threads\": 128, \"ms\": %.4f, \"tflops\": %.2f}\n", \
             NAME, S, bps2[bi], ms, fl / ms * 1e-9);                                          \
    }
    RUN_LDS("lds128_ffma", k_lds_ffma, 1, 0)
    RUN_LDS("lds128_ffma", k_lds_ffma, 2, 0)
    RUN_LDS("lds128_ffma", k_lds_ffma, 4, 0)
    RUN_LDS("lds128_ffma2", k_lds_ffma2, 1, 0)
    RUN_LDS("lds128_ffma2", k_lds_ffma2, 2, 0)
    RUN_LDS("lds128_ffma2", k_lds_ffma2, 4, 0)
    RUN_LDS("lds128_dfma", k_lds_dfma, 1, 0)
    RUN_LDS("lds128_dfma", k_lds_dfma, 2, 0)
  }
  return 0;
}
