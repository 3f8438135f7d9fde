// tc_rollout.cuh — the Blackwell-native contraction path: fused K x T rollout for plain-MLP dynamics
// (nx+nu) -> H1 -> H2 -> nx of ANY hidden widths up to 256 (the host pads them to multiples of 16 with zero units) on tcgen05 tensor cores.
//
// Replaces, like the other rollout kernels, ExperimentalMPPI._compute_total_cost_batch + _compute_rollout_costs
// (tampc/controller/controller.py:340-402) with OnlineMPC._apply_dynamics (online_controller.py:60-79), the network
// forward of NetworkModelWrapper.predict (dynamics/model.py:147-170; the net built by make_sequential_network(h_units=...),
// tampc/util.py:560) and the running / terminal cost (cost.py:11-36,154-189,213-235) inside.
//
// One CTA = one tile of 128 samples = the 128 lanes of tensor memory; one sample per thread for everything that is not a
// matrix product.  Per horizon step:
//   layer 1   D1[128 x NC]  = [x,u] W1^T          tcgen05.mma kind::tf32, M=128, N=NC (32 or 16) per output chunk,
//                                                  A (hi / lo TF32 pieces of the input row) in TMEM, W1 in shared memory
//   epilogue  tcgen05.ld D1 -> LeakyReLU -> hi/lo split -> tcgen05.st: the chunk becomes a K-chunk of layer 2's A operand
//   layer 2   D2[128 x H2] += A1chunk W2chunk^T    N = H2 (up to 256) per instruction, accumulating over the K-chunks;
//                                                  W2 (up to 512 KB of hi/lo pieces) is STREAMED from L2 by the TMA unit
//                                                  (cp.async.bulk + mbarrier complete_tx) through a shared-memory ring of
//                                                  one-K-step stages (16 KB at H2 = 256), or loaded once if it fits
//   layer 3   dx = h2 W3^T + b3                    on the CUDA cores inside the D2 epilogue (nx <= 8 outputs: an N=8 MMA
//                                                  costs as much to issue as an N=256 one; measured ~150 cycles per
//                                                  tcgen05.mma from one thread, tools/peaks/tcgen05_issue.cu)
// 3xTF32: every product is A_lo B_hi + A_hi B_lo + A_hi B_hi with fp32 accumulation in TMEM (the error-compensated split
// of the mma.sync kernels, mma_rollout.cuh); biases enter through a constant ones column of A.
//
// Warp roles (384 threads): warps 0-3 own the samples (state, sampling, cost) and post-process the EVEN layer-1 chunks,
// warps 4-7 post-process the ODD chunks and half of D2 (a warp can only touch the TMEM lanes of its quadrant, warp % 4,
// so two warps share a quadrant); warp 8 issues layer 1, warp 9 issues layer 2 (two issuing threads run in parallel,
// tcgen05_issue.cu; optionally a second one, warp 11, for the other half of its columns), warp 10 feeds the weight ring.  Hand-offs are mbarriers (tcgen05.commit on the MMA side).
// Every wait is bounded: on a timeout the CTA raises an abort flag, all roles leave their loops, TMEM is released and
// the tile's costs come back as NaN.
#pragma once
#include "rollout_kernel.cuh"

namespace tampc {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: 8-row x 16-byte core matrices; LBO = distance between core matrices adjacent in K (128 B: the
// two halves of a K=8 TF32 step lie back to back), SBO = distance between 8-row groups (256 B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
}
// D = f32, A = B = tf32, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(uint32_t N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (SASS UBLKCP), completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// bounded wait; false (and the CTA-wide abort flag raised) on a timeout or when somebody else aborted
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  const long long t0 = clock64();
  while (true) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
    if (*abort_flag || clock64() - t0 > 2000000000ll) {
      *abort_flag = 1;
      return false;
    }
  }
}
// hi piece of the 3xTF32 split ROUNDED to nearest (the mma.sync kernels truncate): |lo| = |f - hi| <= 2^-12 |f|, so the
// tensor core's own truncation of lo to TF32 loses at most 2^-23 |f| — float32 accuracy without a bias towards zero, which
// is what carried a 256-wide network through a horizon of 100 steps at 2e-5 instead of the gate's 1e-5
__device__ __forceinline__ uint32_t split_hi(float f) { return (__float_as_uint(f) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\t.reg .b32 R;\n\telect.sync R|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

#define TC_LD16(r, taddr)                                                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                            \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), \
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                                                    \
               : "r"(taddr)                                                                                                                       \
               : "memory")
#define TC_ST16(taddr, r)                                                                                                                      \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]),     \
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), \
               "r"(r[13]), "r"(r[14]), "r"(r[15])                                                                                                  \
               : "memory")
#define TC_ST8(taddr, r)                                                                                                                                 \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), \
               "r"(r[5]), "r"(r[6]), "r"(r[7])                                                                                                             \
               : "memory")

// mbarrier slots
constexpr int kMaxStages = 32;
enum : int {
  kBarInFull = 0,   // count 256: A_in of the step written (sample owners) and D2 of the previous step read (everybody)
  kBarD1Full = 1,   // [2] count 1 (commit): a layer-1 chunk landed in D1[b]
  kBarA1Full = 3,   // [2] count 128: the group wrote the hi/lo pieces of the chunk into A1[b] (and is done reading D1[b])
  kBarA1Free = 5,   // [2] count 1 (commit): layer 2 consumed the chunk in A1[b] — and the ring stages of its K-steps
                    //     (one commit per chunk serves the epilogue group AND the ring producer: every tcgen05.commit
                    //     costs the tensor pipe ~140 cycles, measured with one commit per K-step)
  kBarD2Full = 7,   // count 1 (commit): layer 2 complete
  kBarWFull = 8,    // [8] count 1 + tx bytes: the K-steps of one K-chunk of W2 landed in the ring (chunk counter % 8; the host
                    //     caps the ring at 7 chunks, so a barrier is never re-armed before its previous phase was consumed).
                    //     The ring itself is handed out one K = 8 step at a time (hi slab + lo slab, 16 KB at H2 = 256), so
                    //     that ~2.75 chunks can be on their way at H2 = 256 instead of one.
  kNumBars = 16
};
constexpr int kWFullBars = 8;
constexpr int kCtrlBytes = ((kNumBars * 8 + 16) + 127) & ~127;
constexpr int kThreads = 384;

// tensor-memory columns (512 allocated: one CTA per SM)
constexpr uint32_t kColD2 = 0, kColD1 = 256, kColA1 = 320, kColAin = 448, kColOnes = 480;

// shared-memory layout behind the image
__host__ __device__ inline int smem_u_off(int image_bytes) { return (image_bytes + 127) & ~127; }
__host__ __device__ inline int smem_ctrl_off(int image_bytes, int T) { return smem_u_off(image_bytes) + (((int)sizeof(float) * T * kMaxNu + 15) & ~15); }
__host__ __device__ inline int smem_dx_off(int image_bytes, int T) { return smem_ctrl_off(image_bytes, T) + kCtrlBytes; }
__host__ __device__ inline int smem_xs_off(int image_bytes, int T) { return smem_dx_off(image_bytes, T) + 128 * 8 * (int)sizeof(float); }
__host__ __device__ inline int smem_ring_off(int image_bytes, int T) { return (smem_xs_off(image_bytes, T) + 128 * 12 * (int)sizeof(float) + 127) & ~127; }

}  // namespace tc

template <int NX, int NU, int NC>
__global__ void __launch_bounds__(tc::kThreads, 1) tc_mlp_rollout_kernel(const RolloutArgs a) {
  using namespace tc;
  static_assert(NC == 32 || NC == 16, "chunk width");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t s_tmem_base;
  const TcMlpDims& D = a.tc;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.src.T;
  CostS<float, float>& cs = *reinterpret_cast<CostS<float, float>*>(smem + D.off_cost);
  SamplerS<float>& sp = *reinterpret_cast<SamplerS<float>*>(smem + D.off_sampler);
  float* Ush = reinterpret_cast<float*>(smem + smem_u_off(a.image_bytes));
  unsigned char* ctrl = smem + smem_ctrl_off(a.image_bytes, T);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ctrl);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(ctrl + kNumBars * 8);
  float* dx_x = reinterpret_cast<float*>(smem + smem_dx_off(a.image_bytes, T));  // [128][8] group 1 -> owners: partial dx
  float* xs_x = reinterpret_cast<float*>(smem + smem_xs_off(a.image_bytes, T));  // [128][12] owners -> group 1: state reached, action seen
  unsigned char* ring = smem + smem_ring_off(a.image_bytes, T);
  auto bar = [&](int i) { return smem_u32(bars + i); };
  const bool prof_cta = a.tc_prof != nullptr && blockIdx.x == 0;
#define TC_STAMP(cond, slot) do { if (prof_cta && (cond)) a.tc_prof[slot] = clock64(); } while (0)

  // ---- prologue: image, nominal sequence, barriers, tensor memory
  grid_dep_launch();
  copy_image_async(smem, a.image, a.image_bytes, tid, kThreads);
  for (int i = tid; i < T * NU; i += kThreads) {
    const int t = i / NU, d = i % NU, ts = t + a.shift;
    Ush[t * kMaxNu + d] = (float)(ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d]);
  }
  if (tid == 0) {
    mbar_init(bar(kBarInFull), 256);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(kBarD1Full + b), 1);
      mbar_init(bar(kBarA1Full + b), 128);
      mbar_init(bar(kBarA1Free + b), D.n_issuers);
    }
    mbar_init(bar(kBarD2Full), D.n_issuers);
    for (int s = 0; s < kWFullBars; ++s) mbar_init(bar(kBarWFull + s), 1);
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  copy_image_wait();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staged weights are read by the tensor core (async proxy)
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = s_tmem_base;
  const int n_chunks = D.n_chunks, k0 = D.k0_steps, h2 = D.h2;

  if (warp < 8) {
    // =================================================================== epilogue groups
    const int grp = warp >> 2, row = tid & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int kbase = blockIdx.x * 128;
    const bool live = kbase + row < a.K_local;
    const int k = live ? kbase + row : a.K_local - 1;
    const float slope = (float)a.slope;
    const float* w3 = reinterpret_cast<const float*>(smem + D.off_w3);  // [h2][2 * ceil(nx/2)]: output pairs of every k
    const float* b3 = reinterpret_cast<const float*>(smem + D.off_b3);
    float x[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = (float)a.state[i];
    double cost = 0.0;
    float pert = 0.f;
    const float bad2 = (float)(a.bad_norm * a.bad_norm);
    uint32_t n_d1 = 0, n_d2 = 0;   // chunks of this group / steps done so far (mbarrier phases)
    bool alive = true;
    if (grp == 0) {
      uint32_t ones[8] = {__float_as_uint(1.0f), __float_as_uint(1.0f), 0u, 0u, 0u, 0u, 0u, 0u};
      TC_ST8(tb + kColOnes + lane_off, ones);
    } else {
      mbar_arrive(bar(kBarInFull));  // nothing of a previous step to finish reading
    }
    // `alive` is warp-uniform (the tcgen05.ld / st below are warp-collective) and turns false when a wait timed out or
    // another role aborted; a step is always walked to its two named barriers, and the second one is a reduction that
    // hands every thread the same stop decision.
    auto wait_all = [&](int b, uint32_t parity) { return __all_sync(0xffffffffu, mbar_wait(bar(b), parity, abort_flag) ? 1 : 0) != 0; };
    // The recurrence x_{t+1} = f(x_t, u_t) is the only thing that must sit between two steps' tensor work, so the sample
    // owners keep everything else off that path: the action of step t+1 is drawn while layer 2 of step t runs, and the
    // running cost of the state reached by step t is evaluated after the input row of step t+1 is on its way.
    float u_next[NU], pert_next = 0.f;
    if (grp == 0) {
      float npost[NU];
      pert_next = sample_action<float, NU>(a.src, sp, Ush, k, 0, u_next, npost);
    }
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      bool bad = false;
      const bool st0 = t == 2 && tid == 0, st1 = t == 2 && tid == 128;
      TC_STAMP(st0, 0);
      if (grp == 0) {
#pragma unroll
        for (int i = 0; i < NU; ++i) u[i] = u_next[i];
        if (a.has_guard) {  // online_controller.py:62-66: the networks (and the cost of this step) see zeros
          float n2 = 0.f;
#pragma unroll
          for (int i = 0; i < NX; ++i) n2 = __fmaf_rn(x[i], x[i], n2);
          bad = n2 > bad2;
          if (bad) {
#pragma unroll
            for (int i = 0; i < NU; ++i) u[i] = 0.f;
          }
        }
        // the input row [x, u, 0...] as hi / lo TF32 pieces: hi in columns [0, 8 k0), lo in [8 k0, 16 k0)
        for (int ks = 0; ks < k0; ++ks) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int f = ks * 8 + j;
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < NX; ++i) v = f == i ? x[i] : v;
#pragma unroll
            for (int i = 0; i < NU; ++i) v = f == NX + i ? u[i] : v;
            v = bad ? 0.f : v;
            hi[j] = split_hi(v);
            lo[j] = __float_as_uint(v - __uint_as_float(hi[j]));
          }
          TC_ST8(tb + kColAin + 8 * ks + lane_off, hi);
          TC_ST8(tb + kColAin + 8 * (k0 + ks) + lane_off, lo);
        }
        wait_st();
        fence_before();
        mbar_arrive(bar(kBarInFull));
        TC_STAMP(st0, 1);
        pert += pert_next;
      }
      // The running cost of the state reached by the previous step is group 1's job (it neither samples nor updates): the
      // owners publish (x_{t}, action seen by step t-1) in shared memory at the end of a step, group 1 evaluates the cost at
      // the top of the next one.  (Measured: on the owners it delayed their own layer-1 chunks, i.e. layer 2.)
      if (grp == 1 && t > 0) {
        const float* r = xs_x + row * 12;
        float xr[NX], ur[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) xr[i] = r[i];
#pragma unroll
        for (int i = 0; i < NU; ++i) ur[i] = r[8 + i];
        cost += (double)running_cost<float, float, NX, NU>(cs, xr, ur);
        if (a.states_out != nullptr && live) {
          double* so = a.states_out + ((size_t)k * T + (t - 1)) * NX;
#pragma unroll
          for (int i = 0; i < NX; ++i) so[i] = (double)xr[i];
        }
      }
      // ---- layer-1 chunks of this group: D1[grp] -> LeakyReLU -> hi/lo -> A1[grp]
      for (int c = grp; c < n_chunks && alive; c += 2) {
        alive = wait_all(kBarD1Full + grp, n_d1 & 1);
        if (!alive) break;
        TC_STAMP(st0 && c == 0, 2);
        TC_STAMP(st1 && c == 1, 48);
        fence_after();
        uint32_t v[NC];
        TC_LD16(v, tb + kColD1 + 32 * grp + lane_off);
        if (NC == 32) TC_LD16((v + 16), tb + kColD1 + 32 * grp + 16 + lane_off);
        wait_ld();
        // A1[grp] may be overwritten once layer 2 has consumed its previous chunk (none before the group's first chunk)
        if (n_d1 > 0) alive = wait_all(kBarA1Free + grp, (n_d1 - 1) & 1);
        if (!alive) break;
        fence_after();
        uint32_t lo[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          float f = __uint_as_float(v[j]);
          f = fmaxf(f, slope * f);
          const uint32_t h = split_hi(f);
          lo[j] = __float_as_uint(f - __uint_as_float(h));
          v[j] = h;
        }
        const uint32_t ta = tb + kColA1 + 64 * grp + lane_off;
        TC_ST16(ta, v);
        if (NC == 32) TC_ST16(ta + 16, (v + 16));
        TC_ST16(ta + NC, lo);
        if (NC == 32) TC_ST16(ta + NC + 16, (lo + 16));
        wait_st();
        fence_before();
        mbar_arrive(bar(kBarA1Full + grp));
        TC_STAMP(st0 && c == 0, 3);
        TC_STAMP(st1 && c == 1, 49);
        ++n_d1;
      }
      TC_STAMP(st0, 4);
      if (grp == 0 && t + 1 < T) {  // the next step's perturbed action, while layer 2 runs
        float npost[NU];
        pert_next = sample_action<float, NU>(a.src, sp, Ush + (t + 1) * kMaxNu, k, t + 1, u_next, npost);
      }
      TC_STAMP(st0, 5);
      // ---- D2 -> LeakyReLU -> layer 3 on the CUDA cores; the two groups take alternate NC-column chunks
      if (alive) alive = wait_all(kBarD2Full, n_d2 & 1);
      ++n_d2;
      TC_STAMP(st0, 6);
      TC_STAMP(st1, 50);
      // layer 3 with packed FFMA2: outputs paired in 64-bit accumulators; W3 lies in shared memory as [k][NXP] float pairs, so
      // two consecutive k are NXP broadcast LDS.128 (the shared-memory pipe, one wavefront per LDS, was half of this phase
      // with one LDS.128 pair per k)
      constexpr int NXP = (NX + 1) / 2;
      uint64_t dxp[NXP];
#pragma unroll
      for (int i = 0; i < NXP; ++i) dxp[i] = grp == 0 ? pack2(b3[2 * i], b3[2 * i + 1]) : pack2(0.f, 0.f);
      if (alive) {
        fence_after();
        for (int c = grp; c < h2 / NC; c += 2) {
          uint32_t v[NC];
          TC_LD16(v, tb + kColD2 + NC * c + lane_off);
          if (NC == 32) TC_LD16((v + 16), tb + kColD2 + NC * c + 16 + lane_off);
          wait_ld();
          const ulonglong2* wrow = reinterpret_cast<const ulonglong2*>(w3 + (size_t)c * NC * 2 * NXP);
#pragma unroll
          for (int j = 0; j < NC; j += 2) {
            float f0 = __uint_as_float(v[j]), f1 = __uint_as_float(v[j + 1]);
            f0 = fmaxf(f0, slope * f0);
            f1 = fmaxf(f1, slope * f1);
            const uint64_t ff[2] = {pack2(f0, f0), pack2(f1, f1)};
            ulonglong2 q[NXP];
#pragma unroll
            for (int i = 0; i < NXP; ++i) q[i] = wrow[(j / 2) * NXP + i];
#pragma unroll
            for (int p = 0; p < 2 * NXP; ++p) {  // pair p: output pair p % NXP of element j + p / NXP
              const uint64_t w = (p & 1) ? q[p / 2].y : q[p / 2].x;
              dxp[p % NXP] = ffma2(ff[p / NXP], w, dxp[p % NXP]);
            }
          }
        }
      }
      float dx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dx[i] = 0.f;
#pragma unroll
      for (int i = 0; i < NXP; ++i) unpack2(dxp[i], dx[2 * i], dx[2 * i + 1]);
      TC_STAMP(st0, 7);
      TC_STAMP(st1, 51);
      if (grp == 1) {
        fence_before();
        mbar_arrive(bar(kBarInFull));  // this thread is done with D2: layer 2 of the next step may overwrite it
        float4* p = reinterpret_cast<float4*>(dx_x + row * 8);
        p[0] = make_float4(dx[0], dx[1], dx[2], dx[3]);
        p[1] = make_float4(dx[4], dx[5], dx[6], dx[7]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      TC_STAMP(st0, 8);
      if (grp == 0) {
        const float4* p = reinterpret_cast<const float4*>(dx_x + row * 8);
        const float4 pa = p[0], pb = p[1];
        const float other[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int i = 0; i < NX; ++i) {
          float xn = x[i] + (dx[i] + other[i]);
          if ((a.wrap_mask >> i) & 1u) xn = angle_normalize<float>(xn);
          x[i] = bad ? 0.f : xn;  // next_state[bad] = state[bad] (= 0), online_controller.py:77
        }
        float* r = xs_x + row * 12;  // for the cost of this step (group 1, next iteration)
#pragma unroll
        for (int i = 0; i < NX; ++i) r[i] = x[i];
#pragma unroll
        for (int i = 0; i < NU; ++i) r[8 + i] = u[i];
      }
      TC_STAMP(st0, 9);
      // the exchange rows are free again before anybody writes the next step's; the barrier also reduces the stop decision
      uint32_t stop;
      const uint32_t mine = alive ? 0u : 1u;
      asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, 2, 256, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(stop) : "r"(mine) : "memory");
      if (stop) break;
    }
    // the running cost of the last step's state (group 1), then the total through the exchange rows
    if (grp == 1) {
      if (T > 0) {
        const float* r = xs_x + row * 12;
        float xr[NX], ur[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) xr[i] = r[i];
#pragma unroll
        for (int i = 0; i < NU; ++i) ur[i] = r[8 + i];
        cost += (double)running_cost<float, float, NX, NU>(cs, xr, ur);
        if (a.states_out != nullptr && live) {
          double* so = a.states_out + ((size_t)k * T + (T - 1)) * NX;
#pragma unroll
          for (int i = 0; i < NX; ++i) so[i] = (double)xr[i];
        }
      }
      *reinterpret_cast<double*>(dx_x + row * 8) = cost;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (grp == 0 && live) {
      double total = *reinterpret_cast<const double*>(dx_x + row * 8) + (double)terminal_cost<float, float, NX>(cs, x) + (double)pert;
      if (*abort_flag) total = __longlong_as_double(0x7ff8000000000000ll);
      a.cost_out[k] = total;
    }
  } else if (warp == 8) {
    // =================================================================== layer-1 issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc(NC);
      const uint32_t w1 = smem_u32(smem + D.off_w1);
      const uint32_t slab = (NC / 8) * 256;
      uint32_t use[2] = {0, 0};
      bool alive = true;
      for (int t = 0; t < T && alive; ++t) {
        if (!(alive = mbar_wait(bar(kBarInFull), t & 1, abort_flag))) break;
        TC_STAMP(t == 2, 32);
        fence_after();
        for (int c = 0; c < n_chunks; ++c) {
          const int b = c & 1;
          // D1[b] is free when the group has read the chunk before the previous one in it (it says so with A1Full)
          if (use[b] > 0 && !(alive = mbar_wait(bar(kBarA1Full + b), (use[b] - 1) & 1, abort_flag))) break;
          fence_after();
          const uint32_t d1 = tb + kColD1 + 32 * b;
          const uint32_t base = w1 + (uint32_t)c * (uint32_t)D.w1_chunk_bytes;
          mma_ts(d1, tb + kColOnes, make_desc(base + 2 * k0 * slab), idesc, 0);  // bias through the ones column
          for (int ks = 0; ks < k0; ++ks) {
            const uint64_t bhi = make_desc(base + ks * slab), blo = make_desc(base + (k0 + ks) * slab);
            const uint32_t ahi = tb + kColAin + 8 * ks, alo = tb + kColAin + 8 * (k0 + ks);
            mma_ts(d1, alo, bhi, idesc, 1);
            mma_ts(d1, ahi, blo, idesc, 1);
            mma_ts(d1, ahi, bhi, idesc, 1);
          }
          commit(bar(kBarD1Full + b));
          TC_STAMP(t == 2 && c < 8, 33 + c);
          ++use[b];
        }
      }
    }
  } else if (warp == 9 || warp == 11) {
    // =================================================================== layer-2 issuer(s)
    // With two issuers each takes one half of the output columns (its own half of D2, of the bias slab and of every
    // weight slab): two threads feed the tensor pipe in parallel (tools/peaks/tcgen05_issue.cu).
    const int qi = warp == 9 ? 0 : 1;
    if (qi < D.n_issuers && elect_one()) {
      const uint32_t n_mine = (uint32_t)(h2 / D.n_issuers);
      const uint32_t idesc = make_idesc(n_mine);
      const uint32_t slab2 = (uint32_t)(h2 / 8) * 256;
      const uint32_t half_off = (uint32_t)qi * (n_mine / 8) * 256;   // this issuer's rows inside a slab
      const uint32_t d2 = tb + kColD2 + (uint32_t)qi * n_mine;
      const uint32_t ring_u = smem_u32(ring);
      const uint32_t stages = (uint32_t)D.stages;
      uint32_t use[2] = {0, 0};
      uint32_t g = 0, gc = 0;  // K = 8 steps / K-chunks consumed so far
      bool alive = true;
      for (int t = 0; t < T && alive; ++t) {
        if (!(alive = mbar_wait(bar(kBarInFull), t & 1, abort_flag))) break;
        TC_STAMP(t == 2 && qi == 0, 16);
        fence_after();
        for (int c = 0; c < n_chunks && alive; ++c, ++gc) {
          const int b = c & 1;
          if (!D.resident || t == 0) {
            if (!(alive = mbar_wait(bar(kBarWFull + (D.resident ? 0 : (gc & 7))), D.resident ? 0u : (gc >> 3) & 1u, abort_flag))) break;
          }
          if (!(alive = mbar_wait(bar(kBarA1Full + b), use[b] & 1, abort_flag))) break;
          fence_after();
          // D2 = bias (through the ones column) right in front of the first chunk: issued at the top of the step it would
          // sit in the tensor pipe ahead of layer 1's first chunk, which everything else of the step waits for
          if (c == 0) mma_ts(d2, tb + kColOnes, make_desc(smem_u32(smem + D.off_b2) + half_off), idesc, 0);
          const uint32_t a1 = tb + kColA1 + 64 * b;
#pragma unroll
          for (int ks = 0; ks < NC / 8; ++ks, ++g) {
            const uint32_t s = D.resident ? (uint32_t)(c * (NC / 8) + ks) : g % stages;
            const uint32_t wbase = ring_u + s * 2 * slab2 + half_off;
            const uint64_t bhi = make_desc(wbase), blo = make_desc(wbase + slab2);
            mma_ts(d2, a1 + NC + 8 * ks, bhi, idesc, 1);
            mma_ts(d2, a1 + 8 * ks, blo, idesc, 1);
            mma_ts(d2, a1 + 8 * ks, bhi, idesc, 1);
          }
          commit(bar(kBarA1Free + b));
          TC_STAMP(t == 2 && qi == 0 && c < 8, 17 + c);
          ++use[b];
        }
        if (alive) commit(bar(kBarD2Full));
        TC_STAMP(t == 2 && qi == 0, 26);
      }
    }
  } else {
    // =================================================================== weight-ring producer (TMA)
    if (elect_one()) {
      const unsigned char* w2 = reinterpret_cast<const unsigned char*>(a.tc_w2);
      const uint32_t ring_u = smem_u32(ring);
      const uint32_t bytes = (uint32_t)D.w2_stage_bytes;  // one K = 8 step: hi slab + lo slab
      const uint32_t n_ksteps = (uint32_t)(D.h1 / 8), stages = (uint32_t)D.stages;
      const uint32_t per_chunk = NC / 8;
      if (D.resident) {
        mbar_expect_tx(bar(kBarWFull), bytes * n_ksteps);
        for (uint32_t q = 0; q < n_ksteps; ++q) tma_load(ring_u + q * bytes, w2 + (size_t)q * bytes, bytes, bar(kBarWFull));
      } else {
        // circular ring of one-K-step stages; stages come back a chunk at a time (kBarA1Free)
        uint32_t g = 0, gc = 0, freed = 0, use[2] = {0, 0};
        int cc = 0;  // next chunk (within its step) whose completion has not been observed yet
        bool alive = true;
        for (int t = 0; t < T && alive; ++t)
          for (int c = 0; c < n_chunks && alive; ++c, ++gc) {
            const uint32_t full = bar(kBarWFull + (gc & 7));
            for (uint32_t ks = 0; ks < per_chunk; ++ks, ++g) {
              while (g >= freed + stages) {
                const int b = cc & 1;
                if (!(alive = mbar_wait(bar(kBarA1Free + b), use[b] & 1, abort_flag))) break;
                ++use[b];
                freed += per_chunk;
                if (++cc == n_chunks) cc = 0;
              }
              if (!alive) break;
              if (ks == 0) mbar_expect_tx(full, bytes * per_chunk);
              tma_load(ring_u + (g % stages) * bytes, w2 + (size_t)(c * per_chunk + ks) * bytes, bytes, full);
            }
          }
      }
    }
  }
  // ---- teardown: every role has left its loop (normally or through the abort flag)
  fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int NX, int NU>
cudaError_t launch_tc_mlp(const RolloutArgs& a, int /*block*/, cudaStream_t stream) {
  const int T = a.src.T;
  const size_t smem = (size_t)tc::smem_ring_off(a.image_bytes, T) + (size_t)a.tc.stages * a.tc.w2_stage_bytes;
  if (smem > 226 * 1024) return cudaErrorInvalidConfiguration;  // (the kernel also has a little static shared memory)
  const int grid = (a.K_local + 127) / 128;
  cudaError_t e;
  if (a.tc.nc == 32) {
    auto kern = tc_mlp_rollout_kernel<NX, NU, 32>;
    static size_t configured = 0;
    if (smem > configured) {
      if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)) != cudaSuccess) return e;
      configured = 226 * 1024;
    }
    kern<<<grid, tc::kThreads, smem, stream>>>(a);
  } else {
    auto kern = tc_mlp_rollout_kernel<NX, NU, 16>;
    static size_t configured = 0;
    if (smem > configured) {
      if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)) != cudaSuccess) return e;
      configured = 226 * 1024;
    }
    kern<<<grid, tc::kThreads, smem, stream>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace tampc
