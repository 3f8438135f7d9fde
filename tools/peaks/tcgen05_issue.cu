// How fast can ONE thread feed tcgen05.mma with small tiles?  (Design input of tc_rollout.cuh: a RexExtract step is ~78 MMAs
// of N = 8..32 per 128-sample tile, so the per-instruction cost, not the tensor rate, decides.)
// Per CTA: `issuers` warps (lane 0 of each) issue `per_commit` MMAs (M=128, N, K=8, kind::tf32, A in TMEM), round-robin
// over `naccs` accumulators, then commit to their own mbarrier and wait.  Reports, per MMA: cycles spent in the issue
// loop (clock64 around it) and cycles until completion; with 1 / 2 / 4 CTAs per SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128) issue_probe(long long* out, int N, int iters, int per_commit, int naccs, int issuers, int tmem_cols) {
  __shared__ __align__(128) float sB[8][64 * 8];  // 8 distinct [<=64 x 8] slabs
  __shared__ __align__(8) unsigned long long mbar[4];
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 8 * 64 * 8; i += 128) (&sB[0][0])[i] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  long long t_issue = 0, t_total = 0;
  uint32_t parity = 0;
  bool ok = true;
  if (warp < issuers) {
    // each issuer: its own accumulators inside [warp * 32 * naccs/issuers ...); A in the last 32 columns
    const uint32_t a_col = tb + tmem_cols - 32;
    const uint32_t d0 = tb + warp * (naccs * 32 / issuers);
    const int my_accs = naccs / issuers;
    for (int it = 0; it < iters && ok; ++it) {
      const long long t0 = clock64();
      if (lane == 0) {
        for (int j = 0; j < per_commit; ++j)
          mma_ts(d0 + (j % my_accs) * 32, a_col + 8 * (j & 3), make_desc(smem_u32(&sB[j & 7][0]), 128, 256), idesc, j >= my_accs);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[warp])) : "memory");
      }
      __syncwarp();
      const long long t1 = clock64();
      uint32_t done = 0;
      long long spins = 0;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar[warp])), "r"(parity) : "memory");
      } while (!done && ++spins < 2000000);
      ok = done != 0;
      parity ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long t2 = clock64();
      t_issue += t1 - t0;
      t_total += t2 - t0;
    }
  }
  if (blockIdx.x == 0 && tid == 0) { out[0] = ok ? t_issue : -1; out[1] = t_total; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(tmem_cols) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int iters = 500;
  for (int ctas : {1, 2, 4})
    for (int N : {8, 16, 32, 64})
      for (int per_commit : {4, 13, 64})
        for (int naccs : {1, 2, 4})
          for (int issuers : {1, 2}) {
            if (naccs < issuers || N * (naccs / issuers) > 64 * 4) continue;
            if (N > 32 && naccs > 2) continue;
            long long h[2] = {0, 0};
            for (int rep = 0; rep < 2; ++rep) {
              issue_probe<<<148 * ctas, 128>>>(d, N, iters, per_commit, naccs, issuers, 512 / ctas < 256 ? 128 : 256);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("{\"bench\": \"tcgen05_issue\", \"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            printf("{\"bench\": \"tcgen05_issue\", \"ctas_per_sm\": %d, \"N\": %d, \"mma_per_commit\": %d, \"accumulators\": %d, \"issuing_warps\": %d, "
                   "\"issue_cycles_per_mma\": %.1f, \"total_cycles_per_mma\": %.1f, \"total_cycles_per_commit\": %.0f}\n",
                   ctas, N, per_commit, naccs, issuers, (double)h[0] / iters / per_commit, (double)h[1] / iters / per_commit, (double)h[1] / iters);
          }
  return 0;
}
