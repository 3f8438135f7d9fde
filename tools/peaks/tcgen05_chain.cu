// What would one 32-wide MLP layer cost on tcgen05?  (Evidence for DESIGN.md §3 "why warp-level mma for 16-32-wide layers".)
//
// One "layer" of a 3xTF32 MLP for a 128-sample tile, the way a tcgen05 rollout kernel would have to chain it:
//   12 x tcgen05.mma.kind::tf32 (M=128, N=32, K=8; A from TMEM, B from shared memory; hi*hi + hi*lo + lo*hi over 4 k-steps)
//   -> tcgen05.commit -> mbarrier wait -> tcgen05.ld of the 128x32 accumulator -> activation / hi-lo split in registers
//   -> 2 x tcgen05.st (A_hi, A_lo for the next layer) -> fences -> next layer.
// Timing only: operand VALUES are arbitrary, so the shared-memory tile layout does not matter; every descriptor field
// is a legal encoding (no-swizzle K-major core matrices).  Reports cycles per layer (clock64) for the full round trip
// and for the MMAs alone, with 1, 2 and 4 tiles (CTAs) per SM.  For comparison, the warp-level path: a 32->32 layer for
// 128 samples is 8 tiles x 48 HMMA x 8 cycles = 3072 tensor cycles over 4 sub-partitions = 768 cycles, no round trip.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;  // descriptor version for sm_100
  return d;         // layout_type = 0: no swizzle
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool FULL>
__global__ void __launch_bounds__(128) layer_chain(long long* cycles, float* sink, int iters) {
  __shared__ __align__(128) float sB[2][4][256];  // [hi/lo][k-step][32 x 8]
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 4 * 256; i += 128) (&sB[0][0][0])[i] = 0.001f * (float)((i * 37) % 19 - 9);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // B tiles visible to the tensor core's proxy
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tmem_base_s;
  const uint32_t tD = tbase, tAhi = tbase + 32, tAlo = tbase + 64;      // column offsets
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;                // this warp's 32 TMEM lanes
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(0.01f * (float)(i + tid % 7));
  auto st32 = [&](uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
        "%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
  };
  st32(tAhi + lane_off);
  st32(tAlo + lane_off);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");

  // instruction descriptor: D = f32, A = B = tf32, K-major both, N = 32, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  uint64_t bdesc[2][4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k) bdesc[h][k] = make_desc(smem_u32(&sB[h][k][0]), 128, 256);

  uint32_t parity = 0;
  long long spins_total = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mma_tf32_ts(tD, tAlo + 8 * k, bdesc[0][k], idesc, k != 0);  // lo * hi
        mma_tf32_ts(tD, tAhi + 8 * k, bdesc[1][k], idesc, 1);       // hi * lo
        mma_tf32_ts(tD, tAhi + 8 * k, bdesc[0][k], idesc, 1);       // hi * hi
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t ok = 0;
    long long spins = 0;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok)
                   : "r"(smem_u32(&mbar)), "r"(parity)
                   : "memory");
    } while (!ok && ++spins < 2000000);
    if (!ok) { if (tid == 0) cycles[1] = -1; break; }  // never hang the box
    spins_total += spins;
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (FULL) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
          "%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
            "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
            "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
            "=r"(r[31])
          : "r"(tD + lane_off)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      uint32_t lo[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {  // leaky ReLU, then the TF32 hi / lo split
        float v = __uint_as_float(r[i]);
        v = fmaxf(v, 0.01f * v) * 0.25f;  // (scaled down so that the chain stays finite)
        const uint32_t hi = __float_as_uint(v) & 0xffffe000u;
        lo[i] = __float_as_uint(v - __uint_as_float(hi));
        r[i] = hi;
      }
      st32(tAhi + lane_off);
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = lo[i];
      st32(tAlo + lane_off);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  const long long t1 = clock64();
  if (blockIdx.x == 0 && tid == 0) {
    if (cycles[1] != -1) cycles[0] = t1 - t0;
    cycles[2] = spins_total;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += __uint_as_float(r[i]);
  if (s == 123.456f) sink[0] = s;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase));
}

template <bool FULL>
void run(const char* name, int ctas_per_sm, long long* d_cycles, float* d_sink) {
  const int iters = 2000;
  long long h[3] = {0, 0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemcpy(d_cycles, h, sizeof h, cudaMemcpyHostToDevice);
    layer_chain<FULL><<<148 * ctas_per_sm, 128>>>(d_cycles, d_sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"bench\": \"tcgen05_layer_chain\", \"variant\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e)); return; }
  }
  long long out[3];
  cudaMemcpy(out, d_cycles, sizeof out, cudaMemcpyDeviceToHost);
  if (out[1] == -1) { printf("{\"bench\": \"tcgen05_layer_chain\", \"variant\": \"%s\", \"error\": \"mbarrier never completed\"}\n", name); return; }
  printf("{\"bench\": \"tcgen05_layer_chain\", \"variant\": \"%s\", \"tiles_per_sm\": %d, \"cycles_per_layer\": %.1f, \"mma_per_layer\": 12, "
         "\"samples_per_tile\": 128, \"wait_spins_per_layer\": %.1f}\n",
         name, ctas_per_sm, (double)out[0] / iters, (double)out[2] / iters);
}

int main() {
  long long* d_cycles;
  float* d_sink;
  cudaMalloc(&d_cycles, 64);
  cudaMalloc(&d_sink, 4);
  for (int c : {1, 2, 4}) {
    run<false>("mma_commit_wait_only", c, d_cycles, d_sink);
    run<true>("mma_ld_split_st_roundtrip", c, d_cycles, d_sink);
  }
  return 0;
}
