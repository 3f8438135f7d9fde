// tcgen05 / TMEM facts the tcgen05 rollout kernel (tampc_b200/csrc/tc_rollout.cuh) is designed around, measured on the box:
//   1. CORRECTNESS of the operand conventions it uses: A (hi / lo TF32 pieces) written to TMEM with tcgen05.st 32x32b by the
//      thread that owns the row, B = weights in shared memory as K-major no-swizzle core matrices (8 rows x 16 bytes), one
//      [N x 8] slab per k-step, D read back with tcgen05.ld 32x32b; checked against a host product for several N, including
//      the N % 16 != 0 shapes, the "ones column" bias trick and accumulation across instructions;
//   2. TMEM load / store bandwidth with 4 / 8 / 16 warps per SM (what bounds the activation + hi/lo-split epilogue);
//   3. dense tcgen05.mma throughput, kind::tf32 and kind::f16, M=128, N = 32 / 64 / 256 (the tensor roofline of bench.py).
// Every wait is bounded: a wrong descriptor must not hang the box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_probe tcgen05_probe.cu && ./tcgen05_probe
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: start address, LBO = byte distance between core matrices adjacent in K, SBO = between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;  // descriptor version (sm_100)
  return d;
}
// kind::tf32 / kind::f16 instruction descriptor: D = f32, A = B = fmt (0 f16, 1 bf16, 2 tf32), both K-major
__device__ __host__ __forceinline__ uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc, bool f16) {
  if (f16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc, bool f16) {
  if (f16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ bool wait_parity(uint32_t mbar, uint32_t parity, long long max_spins = 4000000) {
  uint32_t ok = 0;
  long long spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
  } while (!ok && ++spins < max_spins);
  return ok != 0;
}
#define LD32(r, taddr)                                                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23," \
               "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                                                        \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),          \
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),             \
                 "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),             \
                 "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                                               \
               : "r"(taddr)                                                                                                                      \
               : "memory")
#define ST32(taddr, r)                                                                                                                          \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23," \
               "%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                                                                              \
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),           \
               "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),             \
               "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])  \
               : "memory")
#define LD8(r, taddr)                                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                     \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])             \
               : "r"(taddr)                                                                                                 \
               : "memory")
#define ST8(taddr, r)                                                                                                                        \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), \
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])                                                                                      \
               : "memory")

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// 1. correctness: D[128 x N] = A[128 x K] * W[N x K]^T (+ bias via a ones column), A in TMEM (TS mode), W in smem slabs
// smem W layout: slab ks (k-step of 8) at ks * (N/8) * 256 bytes; inside: [n/8][k/4 (2)][n%8][k%4] floats
__global__ void __launch_bounds__(128) gemm_check(const float* A, const float* W, const float* bias, float* D, int N, int K, int with_bias,
                                                  int* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* sW = reinterpret_cast<float*>(smem);
  const int ksteps = K / 8, slab = (N / 8) * 64;  // floats per slab
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    sW[(k / 8) * slab + (n / 8) * 64 + ((k % 8) / 4) * 32 + (n % 8) * 4 + (k % 4)] = W[i];
  }
  float* sB = sW + ksteps * slab;  // bias slab: [N x 8], column 0 = bias
  for (int i = tid; i < N * 8; i += 128) {
    const int n = i / 8, k = i % 8;
    sB[(n / 8) * 64 + (k / 4) * 32 + (n % 8) * 4 + (k % 4)] = (k == 0 && with_bias) ? bias[n] : 0.f;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(&tbase_s, 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = tbase_s, lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t tD = tb, tA = tb + 256, tOnes = tb + 256 + 64 * 2;  // D: N cols; A: K cols; ones: 8 cols
  // A row of this thread, 8 columns at a time
  for (int c = 0; c < K; c += 8) {
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(A[tid * K + c + j]);
    ST8(tA + c + lane_off, r);
  }
  {
    uint32_t r[8] = {__float_as_uint(1.0f), 0, 0, 0, 0, 0, 0, 0};
    ST8(tOnes + lane_off, r);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  if (tid == 0) {
    const uint32_t idesc = make_idesc(2, 128, N);
    for (int ks = 0; ks < ksteps; ++ks)
      mma_ts(tD, tA + 8 * ks, make_desc(smem_u32(sW + ks * slab), 128, 256), idesc, ks != 0, false);
    if (with_bias) mma_ts(tD, tOnes, make_desc(smem_u32(sB), 128, 256), idesc, 1, false);
    commit(smem_u32(&mbar));
  }
  const bool ok = wait_parity(smem_u32(&mbar), 0);
  fence_after();
  if (!ok) {
    if (tid == 0) *status = -1;
  } else {
    for (int c = 0; c < N; c += 8) {
      uint32_t r[8];
      LD8(r, tD + c + lane_off);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[tid * N + c + j] = __uint_as_float(r[j]);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// 2. TMEM bandwidth: every warp moves 32 lanes x 32 columns (4 KB) per instruction, `iters` times
// mode 0: loads only; 1: stores only; 2: load -> leaky + hi/lo split -> 2 stores (the epilogue of one 32-wide layer)
__global__ void __launch_bounds__(512) tmem_bw(long long* cycles, float* sink, int iters, int mode) {
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tbase_s, 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = tbase_s, lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t col = (uint32_t)((warp >> 2) * 96);  // each warp of a quadrant works on its own 96 columns
  uint32_t r[32], lo[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(0.01f * (float)(i + tid % 7));
  ST32(tb + col + lane_off, r);
  ST32(tb + col + 32 + lane_off, r);
  ST32(tb + col + 64 + lane_off, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
      LD32(r, tb + col + lane_off);
      LD32(lo, tb + col + 32 + lane_off);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else if (mode == 1) {
      ST32(tb + col + lane_off, r);
      ST32(tb + col + 32 + lane_off, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
      LD32(r, tb + col + lane_off);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = __uint_as_float(r[i]);
        v = fmaxf(v, 0.01f * v);
        const uint32_t hi = __float_as_uint(v) & 0xffffe000u;
        lo[i] = __float_as_uint(v - __uint_as_float(hi));
        r[i] = hi;
      }
      ST32(tb + col + 32 + lane_off, r);
      ST32(tb + col + 64 + lane_off, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0) cycles[0] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += __uint_as_float(r[i]) + __uint_as_float(lo[i]);
  if (s == 123.456f) sink[0] = s;
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// 3. dense throughput: one CTA per SM issues `iters` x `per_commit` MMAs (M=128, N, one k-step each), SS or TS
__global__ void __launch_bounds__(128) mma_peak(long long* cycles, int N, int iters, int per_commit, int f16, int ts, int* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* s = reinterpret_cast<float*>(smem);
  const int bytes = (128 + 256) * 32;  // A slab [128 x 32 B] + B slab [256 x 32 B]
  for (int i = tid; i < bytes / 4; i += 128) s[i] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(&tbase_s, 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = tbase_s;
  uint32_t parity = 0;
  const uint32_t idesc = make_idesc(f16 ? 0 : 2, 128, N);
  const uint64_t adesc = make_desc(smem_u32(s), 128, 256), bdesc = make_desc(smem_u32(s + 128 * 8), 128, 256);
  const long long t0 = clock64();
  bool ok = true;
  for (int it = 0; it < iters && ok; ++it) {
    if (tid == 0) {
      for (int j = 0; j < per_commit; ++j) {
        const uint32_t d = tb + (uint32_t)(N <= 128 ? (j & 1) * 256 : 0);  // two accumulators alternate when they fit
        if (ts) mma_ts(d, tb + 504, bdesc, idesc, j > 1, f16 != 0);
        else mma_ss(d, adesc, bdesc, idesc, j > 1, f16 != 0);
      }
      commit(smem_u32(&mbar));
    }
    ok = wait_parity(smem_u32(&mbar), parity);
    parity ^= 1;
    fence_after();
  }
  const long long t1 = clock64();
  if (!ok && tid == 0) *status = -1;
  if (blockIdx.x == 0 && tid == 0) cycles[0] = t1 - t0;
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

static float tf32_round(float v) {  // values used below are exactly representable; kept for clarity
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xffffe000u;
  memcpy(&v, &u, 4);
  return v;
}

static int run_checks(int* d_status, std::initializer_list<int> Ns) {
  // ---- 1. correctness
  for (int with_bias = 0; with_bias <= 1; ++with_bias)
    for (int N : Ns) {
      for (int K : {8, 32, 64}) {
        std::vector<float> A(128 * K), W((size_t)N * K), b(N), D((size_t)128 * N, -777.f), want((size_t)128 * N);
        for (int i = 0; i < 128 * K; ++i) A[i] = tf32_round((float)((i * 37 + 11) % 23 - 11) / 8.f);
        for (int i = 0; i < N * K; ++i) W[i] = tf32_round((float)((i * 53 + 5) % 19 - 9) / 16.f);
        for (int i = 0; i < N; ++i) b[i] = tf32_round((float)(i % 7 - 3) / 4.f);
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double acc = with_bias ? b[n] : 0.0;
            for (int k = 0; k < K; ++k) acc += (double)A[m * K + k] * W[n * K + k];
            want[(size_t)m * N + n] = (float)acc;
          }
        float *dA, *dW, *db, *dD;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
        int st = 0;
        cudaMemcpy(d_status, &st, 4, cudaMemcpyHostToDevice);
        const size_t smem = (size_t)(K / 8 + 1) * (N / 8) * 256 + 256;
        cudaFuncSetAttribute(gemm_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        gemm_check<<<1, 128, smem>>>(dA, dW, db, dD, N, K, with_bias, d_status);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (size_t i = 0; i < D.size(); ++i) maxerr = fmax(maxerr, fabs((double)D[i] - want[i]));
        printf("{\"bench\": \"tcgen05_gemm_check\", \"kind\": \"tf32\", \"M\": 128, \"N\": %d, \"K\": %d, \"bias_via_ones_column\": %d, \"cuda\": \"%s\", "
               "\"completed\": %s, \"max_abs_err\": %.3g}\n",
               N, K, with_bias, cudaGetErrorString(e), st == 0 ? "true" : "false", maxerr);
        cudaFree(dA); cudaFree(dW); cudaFree(db); cudaFree(dD);
        if (e != cudaSuccess) return 1;
      }
    }
  return 0;
}

int main() {
  int* d_status;
  long long* d_cycles;
  float* d_sink;
  cudaMalloc(&d_status, 4);
  cudaMalloc(&d_cycles, 64);
  cudaMalloc(&d_sink, 4);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  // ---- 1. correctness (shapes the PTX ISA lists for M=128: N % 16 == 0)
  if (run_checks(d_status, {16, 32, 64, 128, 256})) return 1;
  // ---- 2. TMEM bandwidth
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16}) {
      const int iters = 4000;
      long long cyc = 0;
      for (int rep = 0; rep < 2; ++rep) {
        tmem_bw<<<148, warps * 32>>>(d_cycles, d_sink, iters, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"bench\": \"tmem_bw\", \"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(&cyc, d_cycles, 8, cudaMemcpyDeviceToHost);
      const double per_iter = (double)cyc / iters;
      const double bytes = mode == 2 ? 4096.0 * 3 : 8192.0;  // per warp per iteration
      printf("{\"bench\": \"tmem_bw\", \"mode\": \"%s\", \"warps_per_sm\": %d, \"cycles_per_iter\": %.1f, \"bytes_per_cycle_per_sm\": %.1f}\n",
             mode == 0 ? "ld 2 x (32 lanes x 32 cols)" : (mode == 1 ? "st 2 x (32x32)" : "ld 32x32 -> leaky + split -> st 2 x (32x32)"), warps, per_iter,
             bytes * warps / per_iter);
    }
  // ---- 3. dense throughput
  for (int f16 = 0; f16 <= 1; ++f16)
    for (int ts = 0; ts <= 1; ++ts)
      for (int N : {16, 32, 64, 256}) {
        const int iters = 200, per_commit = 64;
        int st = 0;
        cudaMemcpy(d_status, &st, 4, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(mma_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        mma_peak<<<148, 128, 16 * 1024>>>(d_cycles, N, 20, per_commit, f16, ts, d_status);
        cudaEventRecord(e0);
        mma_peak<<<148, 128, 16 * 1024>>>(d_cycles, N, iters, per_commit, f16, ts, d_status);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"bench\": \"tcgen05_mma_peak\", \"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        long long cyc = 0;
        cudaMemcpy(&cyc, d_cycles, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost);
        const int Kstep = f16 ? 16 : 8;
        const double flop = 2.0 * 128 * N * Kstep * (double)iters * per_commit * 148;
        printf("{\"bench\": \"tcgen05_mma_peak\", \"kind\": \"%s\", \"a_operand\": \"%s\", \"M\": 128, \"N\": %d, \"K_per_mma\": %d, \"completed\": %s, "
               "\"cycles_per_mma\": %.2f, \"tflops\": %.1f}\n",
               f16 ? "f16" : "tf32", ts ? "tmem" : "smem", N, Kstep, st == 0 ? "true" : "false", (double)cyc / ((double)iters * per_commit), flop / (ms * 1e-3) * 1e-12);
      }
  // ---- 1b. N % 16 != 0 at M=128 (not in the ISA table; last, because an illegal descriptor may kill the context)
  run_checks(d_status, {8, 24, 40});
  return 0;
}
