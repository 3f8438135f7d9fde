// mma_rollout.cuh — warp-level tensor-core variant of the fused K x T rollout.
//
// A warp owns NM m-tiles of samples and evaluates every MLP layer as a chain of warp-level mma.sync tiles whose
// accumulator fragments are fed straight back in as the next layer's A fragments (no shuffles, no shared-memory
// round trip between layers): both m8n8k4.f64 and m16n8k8.tf32 leave lane (g = lane/4, t = lane%4) holding
// features (8n+2t, 8n+2t+1) of its rows, and the host packs the next layer's weights in exactly that k order.
//
//   PolicyF64     mma.sync.m8n8k4.f64 (DMMA): the reference's float64 arithmetic (tampc/controller/controller.py:197)
//                 at the FP64 tensor rate — measured 37.1 TFLOP/s on B200 vs 33.7 for DFMA (tools/peaks), with 8x
//                 fewer issue slots than one-DFMA-per-MAC.
//   PolicyTF32x3  mma.sync.m16n8k8.tf32 three times per tile (A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, fp32 accumulate):
//                 float32-level accuracy at 278/3 = 92.7 TFLOP/s-equivalent (FP32 FMA peak is 74).
//
// Per-sample work (noise, clamp, guard, cost: controller.py:386-395, online_controller.py:60-79, cost.py) stays
// one-thread-per-sample; rows move between the two layouts through a small per-warp shared-memory staging tile.
#pragma once
#include "rollout_kernel.cuh"

namespace tampc {

struct PolicyF64 {
  using T = double;
  static constexpr int MT = 8;  // rows per m-tile
  static constexpr int R = 2;   // accumulator registers per (m-tile, 8-feature block) per lane
  static constexpr int kBlockBytes = 512;
};
struct PolicyTF32x3 {
  using T = float;
  static constexpr int MT = 16;
  static constexpr int R = 4;
  static constexpr int kBlockBytes = 512;
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1688_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// bytes of one packed layer: bias[OUT8*8] then IN8*OUT8 blocks of 32 lanes x 16 B
template <class P> __host__ __device__ constexpr int mma_layer_bytes(int in8, int out8) {
  return out8 * 8 * (int)sizeof(typename P::T) + in8 * out8 * P::kBlockBytes;
}

// out = act(W in + b) for NM m-tiles.  in/out: [m-tile][8-feature block][R] accumulator-layout registers.
template <class P, int NM, int IN8, int OUT8, bool ACT>
__device__ __forceinline__ void mma_layer(const typename P::T (&in)[NM][IN8][P::R], typename P::T (&out)[NM][OUT8][P::R],
                                          const unsigned char* __restrict__ wl, typename P::T slope, int lane) {
  using T = typename P::T;
  const int t = lane & 3;
  const T* bias = reinterpret_cast<const T*>(wl);
  const unsigned char* blocks = wl + OUT8 * 8 * sizeof(T);
#pragma unroll
  for (int o = 0; o < OUT8; ++o) {
    const T b0 = bias[o * 8 + 2 * t], b1 = bias[o * 8 + 2 * t + 1];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      out[m][o][0] = b0;
      out[m][o][1] = b1;
      if constexpr (P::R == 4) { out[m][o][2] = b0; out[m][o][3] = b1; }
    }
  }
#pragma unroll
  for (int n = 0; n < IN8; ++n) {
    if constexpr (P::R == 2) {
#pragma unroll
      for (int o = 0; o < OUT8; ++o) {
        const double2 w = *reinterpret_cast<const double2*>(blocks + ((n * OUT8 + o) * 32 + lane) * 16);
#pragma unroll
        for (int m = 0; m < NM; ++m) {
          dmma884(out[m][o], in[m][n][0], w.x);  // k-step 2n   : features 8n + {0,2,4,6}
          dmma884(out[m][o], in[m][n][1], w.y);  // k-step 2n+1 : features 8n + {1,3,5,7}
        }
      }
    } else {
      uint32_t ahi[NM][4], alo[NM][4];
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        // A fragment order {a0,a1,a2,a3} = {(g,k=t),(g+8,t),(g,t+4),(g+8,t+4)} = accumulator {c0,c2,c1,c3}
        const float a[4] = {in[m][n][0], in[m][n][2], in[m][n][1], in[m][n][3]};
        // split by truncation: hi = top 19 bits (what the tensor core reads anyway), lo = a - hi exactly; the core
        // truncates lo's low bits itself.  (cvt.rna.tf32 has no native SASS on sm_100a: ~5 ALU ops per element.)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ahi[m][i] = __float_as_uint(a[i]) & 0xffffe000u;
          alo[m][i] = __float_as_uint(a[i] - __uint_as_float(ahi[m][i]));
        }
      }
#pragma unroll
      for (int o = 0; o < OUT8; ++o) {
        const uint4 w = *reinterpret_cast<const uint4*>(blocks + ((n * OUT8 + o) * 32 + lane) * 16);  // b0hi b1hi b0lo b1lo
#pragma unroll
        for (int m = 0; m < NM; ++m) {
          mma1688_tf32(out[m][o], alo[m][0], alo[m][1], alo[m][2], alo[m][3], w.x, w.y);
          mma1688_tf32(out[m][o], ahi[m][0], ahi[m][1], ahi[m][2], ahi[m][3], w.z, w.w);
          mma1688_tf32(out[m][o], ahi[m][0], ahi[m][1], ahi[m][2], ahi[m][3], w.x, w.y);
        }
      }
    }
  }
  if (ACT) {
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int o = 0; o < OUT8; ++o)
#pragma unroll
        for (int r = 0; r < P::R; ++r) out[m][o][r] = t_max(out[m][o][r], slope * out[m][o][r]);
  }
}

template <class P, int I8, int A8, int B8, int O8>
struct Mlp3Mma {
  static constexpr int kOff1 = mma_layer_bytes<P>(I8, A8);
  static constexpr int kOff2 = kOff1 + mma_layer_bytes<P>(A8, B8);
  static constexpr int kBytes = kOff2 + mma_layer_bytes<P>(B8, O8);
  template <int NM>
  static __device__ __forceinline__ void eval(const unsigned char* __restrict__ w, const typename P::T (&x)[NM][I8][P::R],
                                              typename P::T (&y)[NM][O8][P::R], typename P::T slope, int lane) {
    typename P::T h1[NM][A8][P::R];
    mma_layer<P, NM, I8, A8, true>(x, h1, w, slope, lane);
    typename P::T h2[NM][B8][P::R];
    mma_layer<P, NM, A8, B8, true>(h1, h2, w + kOff1, slope, lane);
    mma_layer<P, NM, B8, O8, false>(h2, y, w + kOff2, slope, lane);
  }
};

// ---- staging between one-thread-per-sample rows and the accumulator layout --------------------------------
// stage[row][F] of T (F features, padded row stride FS); lane (g,t) of m-tile m holds features (8n+2t, 8n+2t+1)
// of row MT*m+g (and of row MT*m+g+8 for the 16-row TF32 tile).
template <class P, int NM, int NB, int FS>
__device__ __forceinline__ void load_blocks(const typename P::T* __restrict__ stage, typename P::T (&acc)[NM][NB][P::R], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      const typename P::T* p = stage + (P::MT * m + g) * FS + 8 * n + 2 * t;
      acc[m][n][0] = p[0];
      acc[m][n][1] = p[1];
      if constexpr (P::R == 4) { acc[m][n][2] = p[8 * FS]; acc[m][n][3] = p[8 * FS + 1]; }
    }
}
template <class P, int NM, int NB, int FS>
__device__ __forceinline__ void store_blocks(typename P::T* __restrict__ stage, const typename P::T (&acc)[NM][NB][P::R], int col0, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      typename P::T* p = stage + (P::MT * m + g) * FS + col0 + 8 * n + 2 * t;
      p[0] = acc[m][n][0];
      p[1] = acc[m][n][1];
      if constexpr (P::R == 4) { p[8 * FS] = acc[m][n][2]; p[8 * FS + 1] = acc[m][n][3]; }
    }
}

// ---- models ---------------------------------------------------------------------------------------------
// Warp-cooperative predict(): every lane passes its own sample's (x,u) (lanes >= NM*MT pass anything), all lanes
// take part in the mma tiles, and each lane gets its sample's state advanced in place.
template <class P, int NX_, int NU_, int H1, int H2>
struct ModelMlpMma {
  static_assert(NX_ + NU_ <= 8 && NX_ <= 8, "one 8-feature input/output block");
  using T = typename P::T;
  using Net = Mlp3Mma<P, 1, H1 / 8, H2 / 8, 1>;
  static constexpr int NX = NX_, NU = NU_;
  static constexpr int kBytes = Net::kBytes;
  static constexpr int kInStride = 8, kOutStride = 10;  // staging row strides (elements)
  static constexpr int kStageElems = 32 * kInStride + 32 * kOutStride;
  static constexpr int kMacs = (NX_ + NU_) * H1 + H1 * H2 + H2 * NX_;
  template <typename ST, int NM>
  static __device__ __forceinline__ void predict(const unsigned char* __restrict__ w, T slope, T* __restrict__ stage, int lane,
                                                 ST (&x)[NX], const ST (&u)[NU]) {
    T* sin = stage;
    T* sout = stage + 32 * kInStride;
#pragma unroll
    for (int i = 0; i < 8; ++i) sin[lane * kInStride + i] = i < NX ? (T)x[i] : (i < NX + NU ? (T)u[i - NX] : (T)0);
    __syncwarp();
    T in[NM][1][P::R], dx[NM][1][P::R];
    load_blocks<P, NM, 1, kInStride>(sin, in, lane);
    Net::template eval<NM>(w, in, dx, slope, lane);
    store_blocks<P, NM, 1, kOutStride>(sout, dx, 0, lane);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] += (ST)sout[lane * kOutStride + i];
    __syncwarp();
  }
};

template <class P, int NX_, int NU_, int E1, int E2, int NZ, int D1, int D2, int NV, int C1, int C2>
struct ModelRexMma {
  static_assert(NX_ + NU_ <= 8 && NZ <= 8 && NV <= 8 && NX_ * NV <= 32, "block limits");
  using T = typename P::T;
  using Enc = Mlp3Mma<P, 1, E1 / 8, E2 / 8, 1>;
  using Dyn = Mlp3Mma<P, 1, D1 / 8, D2 / 8, 1>;
  static constexpr int CO8 = (NX_ * NV + 7) / 8;
  using Dec = Mlp3Mma<P, 1, C1 / 8, C2 / 8, CO8>;
  static constexpr int NX = NX_, NU = NU_;
  static constexpr int kOffDyn = Enc::kBytes;
  static constexpr int kOffDec = kOffDyn + Dyn::kBytes;
  static constexpr int kOffD0 = kOffDec + Dec::kBytes;
  static constexpr int kBytes = kOffD0 + 8 * (int)sizeof(T);
  static constexpr int kInStride = 8;
  static constexpr int kOutStride = 8 * CO8 + 8 + 2;  // decoder blocks, v block, pad against bank conflicts
  static constexpr int kStageElems = 32 * kInStride + 32 * kOutStride;
  static constexpr int kMacs = (NX_ + NU_) * E1 + E1 * E2 + E2 * NZ + NZ * D1 + D1 * D2 + D2 * NV + NX_ * C1 + C1 * C2 + C2 * NX_ * NV + NX_ * NV;
  template <typename ST, int NM>
  static __device__ __forceinline__ void predict(const unsigned char* __restrict__ w, T slope, T* __restrict__ stage, int lane,
                                                 ST (&x)[NX], const ST (&u)[NU]) {
    T* sin = stage;
    T* sout = stage + 32 * kInStride;
#pragma unroll
    for (int i = 0; i < 8; ++i) sin[lane * kInStride + i] = i < NX ? (T)x[i] : (i < NX + NU ? (T)u[i - NX] : (T)0);
    __syncwarp();
    T in[NM][1][P::R];
    load_blocks<P, NM, 1, kInStride>(sin, in, lane);
    {
      T v[NM][1][P::R];
      {
        T z[NM][1][P::R];
        Enc::template eval<NM>(w, in, z, slope, lane);
        Dyn::template eval<NM>(w + kOffDyn, z, v, slope, lane);
      }
      store_blocks<P, NM, 1, kOutStride>(sout, v, 8 * CO8, lane);
    }
    {
      // the decoder's (extractor-folded) first layer has zero weights on the action features: same input block
      T C[NM][CO8][P::R];
      Dec::template eval<NM>(w + kOffDec, in, C, slope, lane);
      store_blocks<P, NM, CO8, kOutStride>(sout, C, 0, lane);
    }
    __syncwarp();
    const T* row = sout + lane * kOutStride;
    const T* d0 = reinterpret_cast<const T*>(w + kOffD0);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      T dx = d0[i];
#pragma unroll
      for (int j = 0; j < NV; ++j) dx = t_fma(row[i * NV + j], row[8 * CO8 + j], dx);
      x[i] += (ST)dx;
    }
    __syncwarp();
  }
};

// ---- kernel ---------------------------------------------------------------------------------------------
template <class Model, typename ST, typename CT>
struct MmaSmem {
  using T = typename Model::T;
  static constexpr size_t kCost = align16<int>(Model::kBytes);
  static constexpr size_t kSampler = kCost + align16<int>(sizeof(CostS<ST, CT>));
  static constexpr size_t kU = kSampler + align16<int>(sizeof(SamplerS<ST>));
  static __host__ __device__ constexpr size_t stage_off(int T_) { return kU + align16<int>(sizeof(ST) * (size_t)T_ * kMaxNu); }
  static __host__ __device__ constexpr size_t bytes(int T_, int warps) {
    return stage_off(T_) + (size_t)warps * align16<int>(sizeof(T) * Model::kStageElems);
  }
};

template <class Model, typename ST, typename CT, int NM>
__global__ void __launch_bounds__(128) rollout_mma_kernel(const RolloutArgs a) {
  using P_T = typename Model::T;
  constexpr int NX = Model::NX, NU = Model::NU;
  constexpr int SW = NM * (sizeof(P_T) == 8 ? PolicyF64::MT : PolicyTF32x3::MT);  // samples per warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = MmaSmem<Model, ST, CT>;
  unsigned char* w = smem_raw;
  CostS<ST, CT>& cs = *reinterpret_cast<CostS<ST, CT>*>(smem_raw + L::kCost);
  SamplerS<ST>& sp = *reinterpret_cast<SamplerS<ST>*>(smem_raw + L::kSampler);
  ST* Ush = reinterpret_cast<ST*>(smem_raw + L::kU);
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int T = a.src.T;
  P_T* stage = reinterpret_cast<P_T*>(smem_raw + L::stage_off(T) + (size_t)warp * align16<int>(sizeof(P_T) * Model::kStageElems));
  {
    const int4* gw = reinterpret_cast<const int4*>(a.weights);
    int4* sw = reinterpret_cast<int4*>(w);
    for (int i = tid; i < Model::kBytes / 16; i += nthr) sw[i] = gw[i];
    stage_cost<ST, CT>(cs, *a.cost, tid, nthr);
    stage_sampler<ST>(sp, *a.sampler, tid, nthr);
    for (int i = tid; i < T * NU; i += nthr) {
      int t = i / NU, d = i % NU, ts = t + a.shift;
      Ush[t * kMaxNu + d] = (ST)(ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d]);
    }
  }
  __syncthreads();

  const int kk = (blockIdx.x * (nthr >> 5) + warp) * SW + lane;
  const bool live = lane < SW && kk < a.K_local;
  const int k = live ? kk : a.K_local - 1;
  ST x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = (ST)a.state[i];
  ST cost = (ST)0, pert = (ST)0;
  const P_T slope = (P_T)a.slope;
  const ST bad2 = (ST)(a.bad_norm * a.bad_norm);

#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    ST u[NU], npost[NU];
    pert += sample_action<ST, NU>(a.src, sp, Ush + t * kMaxNu, k, t, u, npost);
    bool bad = false;
    if (a.has_guard) {
      ST n2 = (ST)0;
#pragma unroll
      for (int i = 0; i < NX; ++i) n2 = t_fma(x[i], x[i], n2);
      bad = n2 > bad2;
      if (bad) {
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = (ST)0;
#pragma unroll
        for (int i = 0; i < NU; ++i) u[i] = (ST)0;
      }
    }
    Model::template predict<ST, NM>(w, slope, stage, lane, x, u);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      if ((a.wrap_mask >> i) & 1u) x[i] = angle_normalize<ST>(x[i]);
      if (bad) x[i] = (ST)0;
    }
    cost += (ST)running_cost<ST, CT, NX, NU>(cs, x, u);
    if (a.states_out != nullptr && live) {
      double* so = a.states_out + ((size_t)k * T + t) * NX;
#pragma unroll
      for (int i = 0; i < NX; ++i) so[i] = (double)x[i];
    }
  }
  ST total = cost + (ST)terminal_cost<ST, CT, NX>(cs, x);
  total += pert;
  if (live) a.cost_out[k] = (double)total;
}

template <class Model, typename ST, typename CT, int NM>
cudaError_t launch_rollout_mma(const RolloutArgs& a, int block, cudaStream_t stream) {
  auto kern = rollout_mma_kernel<Model, ST, CT, NM>;
  constexpr int SW = NM * (sizeof(typename Model::T) == 8 ? PolicyF64::MT : PolicyTF32x3::MT);
  const int warps = block / 32;
  const size_t smem = MmaSmem<Model, ST, CT>::bytes(a.src.T, warps);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int per_block = warps * SW;
  const int grid = (a.K_local + per_block - 1) / per_block;
  kern<<<grid, block, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace tampc
