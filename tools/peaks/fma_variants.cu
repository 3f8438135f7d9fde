// Operand-delivery variants for the per-sample small-MLP rollout (see fma_peaks.cu).
// "layer" = 32 inputs x 32 outputs, S samples per thread, inputs and outputs in registers.
//  A. weights as constant-bank operands (compiler: LDCU.128 -> UR operands), scalar FFMA
//  B. same, packed FFMA2 (pairs along outputs)
//  C. weights by broadcast LDS.128, inputs in registers, FFMA2
//  D. weights by broadcast LDS.128, inputs re-read from per-thread shared columns (runtime dims)
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

constexpr int NL = 4;            // distinct layers' worth of weights cycled through (4 x 1024 = 16 KB)
constexpr int NW = NL * 1024;
__constant__ float c_w[NW];

// A: constant operands, scalar
template <int S>
__global__ void __launch_bounds__(128) k_const_ffma(float* out, int iters, float x0) {
  float x[S][32], y[S][32];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int i = 0; i < 32; ++i) x[s][i] = x0 * (i + 1) + threadIdx.x * 1e-6f + s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < 32; ++j) y[s][j] = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i)
#pragma unroll
        for (int j = 0; j < 32; ++j)
#pragma unroll
          for (int s = 0; s < S; ++s) y[s][j] = fmaf(x[s][i], c_w[l * 1024 + i * 32 + j], y[s][j]);
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < 32; ++j) x[s][j] = fmaxf(y[s][j], 0.01f * y[s][j]);
    }
  }
  float r = 0;
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < 32; ++j) r += x[s][j];
  if (r == 12345.678f) out[0] = r;
}

// B: constant operands, packed along outputs
template <int S>
__global__ void __launch_bounds__(128) k_const_ffma2(float* out, int iters, float x0) {
  float x[S][32]; uint64_t y[S][16];
  const uint64_t* cw2 = reinterpret_cast<const uint64_t*>(c_w);
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int i = 0; i < 32; ++i) x[s][i] = x0 * (i + 1) + threadIdx.x * 1e-6f + s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < 16; ++j) y[s][j] = 0ull;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        uint64_t xx[S];
#pragma unroll
        for (int s = 0; s < S; ++s) xx[s] = pack2(x[s][i], x[s][i]);
#pragma unroll
        for (int j = 0; j < 16; ++j)
#pragma unroll
          for (int s = 0; s < S; ++s) y[s][j] = ffma2(xx[s], cw2[l * 512 + i * 16 + j], y[s][j]);
      }
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a, b; unpack2(y[s][j], a, b);
          x[s][2 * j] = fmaxf(a, 0.01f * a); x[s][2 * j + 1] = fmaxf(b, 0.01f * b);
        }
    }
  }
  float r = 0;
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < 32; ++j) r += x[s][j];
  if (r == 12345.678f) out[0] = r;
}

// C: LDS.128 broadcast weights, inputs in registers, packed
template <int S>
__global__ void __launch_bounds__(128) k_lds_reg_ffma2(float* out, int iters, float x0) {
  __shared__ ulonglong2 w[NW / 4];
  for (int i = threadIdx.x; i < NW / 4; i += blockDim.x) {
    ulonglong2 v; v.x = pack2(1e-3f * (i % 13), 2e-3f); v.y = pack2(-3e-3f, 4e-3f); w[i] = v;
  }
  __syncthreads();
  float x[S][32]; uint64_t y[S][16];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int i = 0; i < 32; ++i) x[s][i] = x0 * (i + 1) + threadIdx.x * 1e-6f + s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int l = 0; l < NL; ++l) {
      const ulonglong2* wl = w + l * 256;
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < 16; ++j) y[s][j] = 0ull;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        uint64_t xx[S];
#pragma unroll
        for (int s = 0; s < S; ++s) xx[s] = pack2(x[s][i], x[s][i]);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          ulonglong2 ww = wl[i * 8 + j4];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            y[s][2 * j4] = ffma2(xx[s], ww.x, y[s][2 * j4]);
            y[s][2 * j4 + 1] = ffma2(xx[s], ww.y, y[s][2 * j4 + 1]);
          }
        }
      }
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a, b; unpack2(y[s][j], a, b);
          x[s][2 * j] = fmaxf(a, 0.01f * a); x[s][2 * j + 1] = fmaxf(b, 0.01f * b);
        }
    }
  }
  float r = 0;
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < 32; ++j) r += x[s][j];
  if (r == 12345.678f) out[0] = r;
}

// D: LDS.128 broadcast weights, inputs in per-thread shared columns [i][thread] of float4 (S=4)
//    or float2 (S=2); runtime in_dim; outputs (32 x S) accumulate in registers and overwrite the column.
template <int S, bool PACKED>
__global__ void __launch_bounds__(128) k_lds_smem(float* out, int iters, float x0, int in_dim) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ulonglong2* w = reinterpret_cast<ulonglong2*>(smem_raw);              // NW floats
  float* act = reinterpret_cast<float*>(smem_raw + NW * 4);            // [32][128][S]
  for (int i = threadIdx.x; i < NW / 4; i += blockDim.x) {
    ulonglong2 v; v.x = pack2(1e-3f * (i % 13), 2e-3f); v.y = pack2(-3e-3f, 4e-3f); w[i] = v;
  }
  for (int i = 0; i < 32; ++i)
#pragma unroll
    for (int s = 0; s < S; ++s) act[(i * 128 + threadIdx.x) * S + s] = x0 * (i + 1) + threadIdx.x * 1e-6f + s;
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int l = 0; l < NL; ++l) {
      const ulonglong2* wl = w + l * 256;
      if (PACKED) {
        uint64_t y[S][16];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int j = 0; j < 16; ++j) y[s][j] = 0ull;
#pragma unroll 2
        for (int i = 0; i < in_dim; ++i) {
          float xv[S];
          if (S == 4) { float4 t = *reinterpret_cast<const float4*>(&act[(i * 128 + threadIdx.x) * 4]); xv[0] = t.x; xv[1] = t.y; xv[2 % S] = t.z; xv[3 % S] = t.w; }
          else if (S == 2) { float2 t = *reinterpret_cast<const float2*>(&act[(i * 128 + threadIdx.x) * 2]); xv[0] = t.x; xv[1] = t.y; }
          else xv[0] = act[i * 128 + threadIdx.x];
          uint64_t xx[S];
#pragma unroll
          for (int s = 0; s < S; ++s) xx[s] = pack2(xv[s], xv[s]);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            ulonglong2 ww = wl[i * 8 + j4];
#pragma unroll
            for (int s = 0; s < S; ++s) {
              y[s][2 * j4] = ffma2(xx[s], ww.x, y[s][2 * j4]);
              y[s][2 * j4 + 1] = ffma2(xx[s], ww.y, y[s][2 * j4 + 1]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float o[2][S];
#pragma unroll
          for (int s = 0; s < S; ++s) { float a, b; unpack2(y[s][j], a, b); o[0][s] = fmaxf(a, 0.01f * a); o[1][s] = fmaxf(b, 0.01f * b); }
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int s = 0; s < S; ++s) act[((2 * j + h) * 128 + threadIdx.x) * S + s] = o[h][s];
        }
      } else {
        float y[S][32];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
          for (int j = 0; j < 32; ++j) y[s][j] = 0.f;
        const float4* wf = reinterpret_cast<const float4*>(wl);
#pragma unroll 2
        for (int i = 0; i < in_dim; ++i) {
          float xv[S];
          if (S == 4) { float4 t = *reinterpret_cast<const float4*>(&act[(i * 128 + threadIdx.x) * 4]); xv[0] = t.x; xv[1] = t.y; xv[2 % S] = t.z; xv[3 % S] = t.w; }
          else if (S == 2) { float2 t = *reinterpret_cast<const float2*>(&act[(i * 128 + threadIdx.x) * 2]); xv[0] = t.x; xv[1] = t.y; }
          else xv[0] = act[i * 128 + threadIdx.x];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 ww = wf[i * 8 + j4];
#pragma unroll
            for (int s = 0; s < S; ++s) {
              y[s][4 * j4 + 0] = fmaf(xv[s], ww.x, y[s][4 * j4 + 0]);
              y[s][4 * j4 + 1] = fmaf(xv[s], ww.y, y[s][4 * j4 + 1]);
              y[s][4 * j4 + 2] = fmaf(xv[s], ww.z, y[s][4 * j4 + 2]);
              y[s][4 * j4 + 3] = fmaf(xv[s], ww.w, y[s][4 * j4 + 3]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j)
#pragma unroll
          for (int s = 0; s < S; ++s) act[(j * 128 + threadIdx.x) * S + s] = fmaxf(y[s][j], 0.01f * y[s][j]);
      }
    }
  }
  float r = 0;
  for (int j = 0; j < 32; ++j)
#pragma unroll
    for (int s = 0; s < S; ++s) r += act[(j * 128 + threadIdx.x) * S + s];
  if (r == 12345.678f) out[0] = r;
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
  std::vector<float> hw(NW); for (int i = 0; i < NW; ++i) hw[i] = 1e-3f * ((i % 7) - 3);
  CK(cudaMemcpyToSymbol(c_w, hw.data(), NW * 4));
  const int iters = 100;
#define REPORT(NAME, S, BPS, MS) printf("{\"bench\": \"%s\", \"samples_per_thread\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", \
    NAME, S, BPS, MS, 2.0 * sms * BPS * 128.0 * iters * (double)NW * S / MS * 1e-9)
  for (int bps : {1, 2, 3, 4, 6}) {
    int grid = sms * bps;
    { double ms = time_ms([&] { k_const_ffma<1><<<grid, 128>>>(out, iters, 1e-3f); }); REPORT("const_ffma", 1, bps, ms); }
    { double ms = time_ms([&] { k_const_ffma<2><<<grid, 128>>>(out, iters, 1e-3f); }); REPORT("const_ffma", 2, bps, ms); }
    { double ms = time_ms([&] { k_const_ffma2<1><<<grid, 128>>>(out, iters, 1e-3f); }); REPORT("const_ffma2", 1, bps, ms); }
    { double ms = time_ms([&] { k_const_ffma2<2><<<grid, 128>>>(out, iters, 1e-3f); }); REPORT("const_ffma2", 2, bps, ms); }
    { double ms = time_ms([&] { k_lds_reg_ffma2<1><<<grid, 128>>>(out, iters, 1e-3f); }); REPORT("lds_reg_ffma2", 1, bps, ms); }
    { double ms = time_ms([&] { k_lds_reg_ffma2<2><<<grid, 128>>>(out, iters, 1e-3f); }); REPORT("lds_reg_ffma2", 2, bps, ms); }
#define RUN_D(S, PACKED, NAME) { \
      size_t sh = NW * 4 + 32 * 128 * S * 4; \
      CK(cudaFuncSetAttribute(k_lds_smem<S, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh)); \
      int nb = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_lds_smem<S, PACKED>, 128, sh)); \
      if (bps <= nb) { double ms = time_ms([&] { k_lds_smem<S, PACKED><<<grid, 128, sh>>>(out, iters, 1e-3f, 32); }); REPORT(NAME, S, bps, ms); } }
    RUN_D(2, true, "lds_smem_ffma2") RUN_D(4, true, "lds_smem_ffma2")
    RUN_D(2, false, "lds_smem_ffma") RUN_D(4, false, "lds_smem_ffma")
  }
  return 0;
}
