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
// With fewer than 32 samples per warp (small K: latency-bound) the spare lanes mirror a sample and take a share
// of its trap-set cost terms, and the horizon's standard normals are drawn up front by all 32 lanes.
#pragma once
#include "rollout_kernel.cuh"
#include "fused_reduce.cuh"
#include <type_traits>

namespace tampc {

// Section timing for development (-DTAMPC_PROFILE): per-warp cycle counts summed into g_prof[section].
#ifdef TAMPC_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF_T0() long long prof_t = clock64()
#define PROF_RESET() prof_t = clock64()
#define PROF(sec) do { long long now_ = clock64(); if ((threadIdx.x & 31) == 0) atomicAdd(&g_prof[sec], (unsigned long long)(now_ - prof_t)); prof_t = clock64(); } while (0)
#else
#define PROF_T0() do {} while (0)
#define PROF_RESET() do {} while (0)
#define PROF(sec) do {} while (0)
#endif

struct PolicyF64 {
  using T = double;
  static constexpr int MT = 8;  // rows per m-tile
  static constexpr int R = 2;   // accumulator registers per (m-tile, 8-feature block) per lane
  static constexpr int PROD = 2;  // mma instructions per (8-feature input block, output block)
  static constexpr int kBlockBytes = 512;
};
struct PolicyTF32x3 {
  using T = float;
  static constexpr int MT = 16;
  static constexpr int R = 4;
  static constexpr int PROD = 3;
  static constexpr int kBlockBytes = 512;
};

// 3xFP16: the same error-compensated split on mma.sync.m16n8k16.f16 — fp16 carries the same 11 significant bits as TF32
// but one instruction covers k = 16, so a layer needs half the tensor instructions (192 -> 108 per 16-sample step for the
// RexExtract shapes).  The low parts would fall into fp16's subnormal range (|lo| <= 2^-11 |a|), so both A_lo and B_lo
// are carried scaled by 2^11 and their products go to a second accumulator that is folded in with 2^-11 at the end:
//   out = bias + sum A_hi B_hi + 2^-11 sum (A_lo' B_hi + A_hi B_lo').
// High parts saturate at fp16's largest finite value (65504): inputs beyond it — only diverged rollouts just below the
// 1e5 bad-state guard — lose accuracy instead of turning into NaN.
struct PolicyF16x3 {
  using T = float;
  static constexpr int MT = 16;
  static constexpr int R = 4;
  static constexpr int PROD = 3;
  static constexpr int kBlockBytes = 512;  // one (k16, n8) block: 32 lanes x {b0hi, b1hi, b0lo', b1lo'} half2 pairs
};
constexpr float kF16LoScale = 2048.f, kF16LoUnscale = 1.f / 2048.f;

__device__ __forceinline__ void mma16816_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1, float c2, float c3) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
// {x -> low half, y -> high half}, round to nearest, saturating
__device__ __forceinline__ uint32_t pack_half2_sat(float x, float y) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
  return r;
}
// (hi, lo') of the pair (x, y): hi = the top 11 significant bits (exact in fp16 within its range), lo' = (v - hi) * 2^11
__device__ __forceinline__ void split_half2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const float hx = __uint_as_float(__float_as_uint(x) & 0xffffe000u), hy = __uint_as_float(__float_as_uint(y) & 0xffffe000u);
  hi = pack_half2_sat(hx, hy);
  lo = pack_half2_sat((x - hx) * kF16LoScale, (y - hy) * kF16LoScale);
}

// d = a*b + c with separate c (first product of an accumulation chain: c = bias or zero)
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b, double c0, double c1) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d[0]), "=d"(d[1]) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__device__ __forceinline__ void mma1688_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1, float c2, float c3) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}

// bytes of one packed layer: bias[OUT8*8] then IN8*OUT8 blocks of 32 lanes x 16 B
template <class P> __host__ __device__ constexpr int mma_layer_bytes(int in8, int out8) {
  return out8 * 8 * (int)sizeof(typename P::T) + in8 * out8 * P::kBlockBytes;
}
// fp16: one block per (16 input features, 8 outputs); an odd number of 8-feature input blocks is zero-padded
template <> __host__ __device__ constexpr int mma_layer_bytes<PolicyF16x3>(int in8, int out8) {
  return out8 * 8 * (int)sizeof(float) + ((in8 + 1) / 2) * out8 * PolicyF16x3::kBlockBytes;
}

// 3xFP16 layer: in / out in the accumulator layout (lane (g,t): features 8n+2t, 8n+2t+1 of rows g and g+8), which
// is also the m16n8k16 A-fragment layout of two consecutive 8-feature blocks — no data movement between layers.
template <int NM, int IN8, int OUT8, bool ACT>
__device__ __forceinline__ void mma_layer_f16(const float (&in)[NM][IN8][4], float (&out)[NM][OUT8][4], const unsigned char* __restrict__ wl,
                                              float slope, int lane) {
  constexpr int IN16 = (IN8 + 1) / 2;
  const int t = lane & 3;
  const float* bias = reinterpret_cast<const float*>(wl);
  const unsigned char* blocks = wl + OUT8 * 8 * sizeof(float);
  float accm[NM][OUT8][4], accx[NM][OUT8][4];
  float b0[OUT8], b1[OUT8];
#pragma unroll
  for (int o = 0; o < OUT8; ++o) { b0[o] = bias[o * 8 + 2 * t]; b1[o] = bias[o * 8 + 2 * t + 1]; }
#pragma unroll
  for (int k = 0; k < IN16; ++k) {
    uint32_t ahi[NM][4], alo[NM][4];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      // a0 = (row g; k 2t,2t+1)  a1 = (row g+8; same k)  a2 = (row g; k 2t+8,2t+9)  a3 = (row g+8; same)
      split_half2(in[m][2 * k][0], in[m][2 * k][1], ahi[m][0], alo[m][0]);
      split_half2(in[m][2 * k][2], in[m][2 * k][3], ahi[m][1], alo[m][1]);
      if (2 * k + 1 < IN8) {
        split_half2(in[m][2 * k + 1][0], in[m][2 * k + 1][1], ahi[m][2], alo[m][2]);
        split_half2(in[m][2 * k + 1][2], in[m][2 * k + 1][3], ahi[m][3], alo[m][3]);
      } else {
        ahi[m][2] = ahi[m][3] = alo[m][2] = alo[m][3] = 0u;
      }
    }
#pragma unroll
    for (int o = 0; o < OUT8; ++o) {
      const uint4 w = *reinterpret_cast<const uint4*>(blocks + ((k * OUT8 + o) * 32 + lane) * 16);  // b0hi b1hi b0lo' b1lo'
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        if (k == 0) {
          mma16816_f16(accm[m][o], ahi[m], w.x, w.y, b0[o], b1[o], b0[o], b1[o]);
          mma16816_f16(accx[m][o], alo[m], w.x, w.y, 0.f, 0.f, 0.f, 0.f);
        } else {
          mma16816_f16(accm[m][o], ahi[m], w.x, w.y, accm[m][o][0], accm[m][o][1], accm[m][o][2], accm[m][o][3]);
          mma16816_f16(accx[m][o], alo[m], w.x, w.y, accx[m][o][0], accx[m][o][1], accx[m][o][2], accx[m][o][3]);
        }
        mma16816_f16(accx[m][o], ahi[m], w.z, w.w, accx[m][o][0], accx[m][o][1], accx[m][o][2], accx[m][o][3]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int o = 0; o < OUT8; ++o)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float v = __fmaf_rn(accx[m][o][r], kF16LoUnscale, accm[m][o][r]);
        out[m][o][r] = ACT ? fmaxf(v, slope * v) : v;
      }
}

// out = act(W in + b) for NM m-tiles.  in/out: [m-tile][8-feature block][R] accumulator-layout registers.
// Narrow outputs (OUT8 < 4) would be one long dependent mma chain per tile; their products are dealt over NS
// independent accumulators that are added at the end.
template <class P, int NM, int IN8, int OUT8, bool ACT>
__device__ __forceinline__ void mma_layer(const typename P::T (&in)[NM][IN8][P::R], typename P::T (&out)[NM][OUT8][P::R],
                                          const unsigned char* __restrict__ wl, typename P::T slope, int lane) {
  using T = typename P::T;
  if constexpr (std::is_same<P, PolicyF16x3>::value) {
    mma_layer_f16<NM, IN8, OUT8, ACT>(in, out, wl, slope, lane);
    return;
  }
  constexpr int NPROD = P::PROD * IN8;                       // products accumulated per output tile
  constexpr int NS_WANT = OUT8 >= 4 ? 1 : (OUT8 == 2 ? 2 : 4);
  constexpr int NS = NS_WANT < NPROD ? NS_WANT : NPROD;      // accumulators per output tile
  const int t = lane & 3;
  const T* bias = reinterpret_cast<const T*>(wl);
  const unsigned char* blocks = wl + OUT8 * 8 * sizeof(T);
  T acc[NS][NM][OUT8][P::R];
  T b0[OUT8], b1[OUT8];
#pragma unroll
  for (int o = 0; o < OUT8; ++o) { b0[o] = bias[o * 8 + 2 * t]; b1[o] = bias[o * 8 + 2 * t + 1]; }
#pragma unroll
  for (int n = 0; n < IN8; ++n) {
    if constexpr (P::R == 2) {
#pragma unroll
      for (int o = 0; o < OUT8; ++o) {
        const double2 w = *reinterpret_cast<const double2*>(blocks + ((n * OUT8 + o) * 32 + lane) * 16);
#pragma unroll
        for (int m = 0; m < NM; ++m) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {  // k-step 2n+j: features 8n + {j, j+2, j+4, j+6}
            const int q = 2 * n + j, s = q % NS;
            const double a = in[m][n][j], b = j ? w.y : w.x;
            if (q < NS) dmma884(acc[s][m][o], a, b, s == 0 ? b0[o] : 0.0, s == 0 ? b1[o] : 0.0);
            else dmma884(acc[s][m][o], a, b, acc[s][m][o][0], acc[s][m][o][1]);
          }
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
#pragma unroll
          for (int j = 0; j < 3; ++j) {  // small terms first: A_lo*B_hi, A_hi*B_lo, A_hi*B_hi
            const int q = 3 * n + j, s = q % NS;
            const uint32_t wb0 = j == 1 ? w.z : w.x, wb1 = j == 1 ? w.w : w.y;
            const uint32_t(&af)[4] = j == 0 ? alo[m] : ahi[m];
            if (q < NS) {
              const float c0 = s == 0 ? b0[o] : 0.f, c1 = s == 0 ? b1[o] : 0.f;
              mma1688_tf32(acc[s][m][o], af, wb0, wb1, c0, c1, c0, c1);
            } else {
              mma1688_tf32(acc[s][m][o], af, wb0, wb1, acc[s][m][o][0], acc[s][m][o][1], acc[s][m][o][2], acc[s][m][o][3]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int o = 0; o < OUT8; ++o)
#pragma unroll
      for (int r = 0; r < P::R; ++r) {
        T v = acc[0][m][o][r];
#pragma unroll
        for (int s = 1; s < NS; ++s) v += acc[s][m][o][r];
        out[m][o][r] = ACT ? t_max(v, slope * v) : v;
      }
}

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

// three-layer MLP layout  I8 -> A8 -> B8 -> O8 (in 8-feature blocks)
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

// ---- models ---------------------------------------------------------------------------------------------
// Warp-cooperative predict(): every lane passes the (x,u) of the sample row it holds (rows >= NM*MT are mirrors or
// don't-cares), all lanes take part in the mma tiles, and each lane gets its row's state advanced in place.
template <class P, int NX_, int NU_, int H1, int H2>
struct ModelMlpMma {
  static_assert(NX_ + NU_ <= 8 && NX_ <= 8, "one 8-feature input/output block");
  using T = typename P::T;
  using Policy = P;
  using Net = Mlp3Mma<P, 1, H1 / 8, H2 / 8, 1>;
  static constexpr int NX = NX_, NU = NU_;
  static constexpr int kBytes = Net::kBytes;
  static constexpr int kInStride = 8, kOutStride = 10;  // staging row strides (elements)
  static constexpr int kStageElems = 32 * kInStride + 32 * kOutStride;
  static constexpr int kMacs = (NX_ + NU_) * H1 + H1 * H2 + H2 * NX_;
  template <typename ST, int NM>
  static __device__ __forceinline__ void predict(const unsigned char* __restrict__ w, T slope, T* __restrict__ stage, int lane, int row,
                                                 ST (&x)[NX], const ST (&u)[NU]) {
    T* sin = stage;
    T* sout = stage + 32 * kInStride;
    if (lane == row) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sin[row * kInStride + i] = i < NX ? (T)x[i] : (i < NX + NU ? (T)u[i - NX] : (T)0);
    }
    __syncwarp();
    T in[NM][1][P::R], dx[NM][1][P::R];
    load_blocks<P, NM, 1, kInStride>(sin, in, lane);
    Net::template eval<NM>(w, in, dx, slope, lane);
    store_blocks<P, NM, 1, kOutStride>(sout, dx, 0, lane);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] += (ST)sout[row * kOutStride + i];
    __syncwarp();
  }
};

template <class P, int NX_, int NU_, int E1, int E2, int NZ, int D1, int D2, int NV, int C1, int C2>
struct ModelRexMma {
  static_assert(NX_ + NU_ <= 8 && NZ <= 8 && NV <= 8 && NX_ * NV <= 32, "block limits");
  using T = typename P::T;
  using Policy = P;
  using Enc = Mlp3Mma<P, 1, E1 / 8, E2 / 8, 1>;
  using Dyn = Mlp3Mma<P, 1, D1 / 8, D2 / 8, 1>;
  static constexpr int CO8 = (NX_ * NV + 7) / 8;
  static constexpr int kNV = NV;
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
  static __device__ __forceinline__ void predict(const unsigned char* __restrict__ w, T slope, T* __restrict__ stage, int lane, int row,
                                                 ST (&x)[NX], const ST (&u)[NU]) {
    T* sin = stage;
    T* sout = stage + 32 * kInStride;
    PROF_T0();
    if (lane == row) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sin[row * kInStride + i] = i < NX ? (T)x[i] : (i < NX + NU ? (T)u[i - NX] : (T)0);
    }
    __syncwarp();
    T in[NM][1][P::R];
    load_blocks<P, NM, 1, kInStride>(sin, in, lane);
    PROF(2);
    {
      // The encoder->latent-dynamics chain and the decoder are independent until dx = C v: their layers are
      // issued alternately so that the tensor pipe always has a second dependency chain to run.
      // (The decoder's extractor-folded first layer has zero weights on the action features: same input block.)
      T e1[NM][E1 / 8][P::R], c1[NM][C1 / 8][P::R];
      mma_layer<P, NM, 1, E1 / 8, true>(in, e1, w, slope, lane);
      mma_layer<P, NM, 1, C1 / 8, true>(in, c1, w + kOffDec, slope, lane);
      T e2[NM][E2 / 8][P::R], c2[NM][C2 / 8][P::R];
      mma_layer<P, NM, E1 / 8, E2 / 8, true>(e1, e2, w + Enc::kOff1, slope, lane);
      mma_layer<P, NM, C1 / 8, C2 / 8, true>(c1, c2, w + kOffDec + Dec::kOff1, slope, lane);
      T z[NM][1][P::R];
      mma_layer<P, NM, E2 / 8, 1, false>(e2, z, w + Enc::kOff2, slope, lane);
      {
        T C[NM][CO8][P::R];
        mma_layer<P, NM, C2 / 8, CO8, false>(c2, C, w + kOffDec + Dec::kOff2, slope, lane);
        store_blocks<P, NM, CO8, kOutStride>(sout, C, 0, lane);
      }
      PROF(3);
      T v[NM][1][P::R];
      Dyn::template eval<NM>(w + kOffDyn, z, v, slope, lane);
      store_blocks<P, NM, 1, kOutStride>(sout, v, 8 * CO8, lane);
      PROF(4);
    }
    __syncwarp();
    const T* r = sout + row * kOutStride;
    const T* d0 = reinterpret_cast<const T*>(w + kOffD0);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      T dx = d0[i];
#pragma unroll
      for (int j = 0; j < NV; ++j) dx = t_fma(r[i * NV + j], r[8 * CO8 + j], dx);
      x[i] += (ST)dx;
    }
    __syncwarp();
    PROF(6);
  }
};

// ---- kernel ---------------------------------------------------------------------------------------------
template <class Model, typename ST, typename CT>
struct MmaSmem {
  using T = typename Model::T;
  static constexpr size_t kCost = align16<int>(Model::kBytes);
  static constexpr size_t kSampler = kCost + align16<int>(sizeof(CostS<ST, CT>));
  static constexpr size_t kSamplerD = kSampler + align16<int>(sizeof(SamplerS<ST>));  // float64 copy for the fused update
  static constexpr size_t kU = kSamplerD + align16<int>(sizeof(SamplerS<double>));
  static __host__ __device__ constexpr size_t stage_off(int T_) { return kU + align16<int>(sizeof(ST) * (size_t)T_ * kMaxNu); }
  static __host__ __device__ constexpr size_t warp_bytes(int T_, int sw, bool pre) {
    return align16<int>(sizeof(T) * Model::kStageElems) + (pre ? align16<int>(sizeof(float) * (size_t)T_ * sw * Model::NU) : 0);
  }
  static __host__ __device__ constexpr size_t bytes(int T_, int warps, int sw, bool pre) { return stage_off(T_) + (size_t)warps * warp_bytes(T_, sw, pre); }
};

// PRE: draw the whole horizon's standard normals into shared memory before the loop (all 32 lanes, independent
// Philox blocks -> instruction-level parallelism), instead of one dependent Philox + Box-Muller chain per step.
// BLK: largest block the instantiation is launched with.  256 asks for two resident CTAs (<= 128 registers): eight
// warps share one copy of the image, 16 warps per SM instead of the 12 that three 4-warp CTAs give.
// Everything one warp does after the image is staged: SW = 16 * NM samples starting at kbase (8 * NM for float64).
// warp_index / warps_total: position in the grid-wide list of warps (fused tail only).
template <class Model, typename ST, typename CT, int NM, bool PRE>
__device__ __forceinline__ void rollout_mma_warp(const RolloutArgs& a, unsigned char* smem_raw, unsigned char* wbase, int kbase, int lane,
                                                 int warp_index, int warps_total) {
  using P_T = typename Model::T;
  constexpr int NX = Model::NX, NU = Model::NU;
  constexpr int SW = NM * (sizeof(P_T) == 8 ? PolicyF64::MT : PolicyTF32x3::MT);  // samples per warp
  constexpr int NPARTS = 32 / SW;                                                  // lanes holding the same sample
  using L = MmaSmem<Model, ST, CT>;
  unsigned char* w = smem_raw;
  CostS<ST, CT>& cs = *reinterpret_cast<CostS<ST, CT>*>(smem_raw + L::kCost);
  SamplerS<ST>& sp = *reinterpret_cast<SamplerS<ST>*>(smem_raw + L::kSampler);
  ST* Ush = reinterpret_cast<ST*>(smem_raw + L::kU);
  const int T = a.src.T;
  P_T* stage = reinterpret_cast<P_T*>(wbase);
  float* zpre = reinterpret_cast<float*>(wbase + align16<int>(sizeof(P_T) * Model::kStageElems));  // [T][SW][NU]
  const int row = lane % SW, part = lane / SW;
  const bool live = kbase + row < a.K_local;
  const int k = live ? kbase + row : a.K_local - 1;
  const bool use_pre = PRE && a.src.mode == kModePhilox;
  if (use_pre) {
#pragma unroll 2
    for (int item = lane; item < T * SW; item += 32) {
      const int t = item / SW, r = item % SW;
      const int kr = min(kbase + r, a.K_local - 1);
      float z[4];
      philox_normal4(a.src.seed, a.src.call, (uint32_t)(a.src.sample_offset + kr), (uint32_t)t, z);
#pragma unroll
      for (int i = 0; i < NU; ++i) zpre[item * NU + i] = z[i];
    }
    __syncwarp();  // the draws are private to the warp
  }

  ST x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = (ST)a.state[i];
  double cost = 0.0;  // running-cost accumulator: float64 in every mode (one add per step)
  ST pert = (ST)0;
  const P_T slope = (P_T)a.slope;
  const ST bad2 = (ST)(a.bad_norm * a.bad_norm);

#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    ST u[NU], npost[NU];
    PROF_T0();
    pert += sample_action<ST, NU>(a.src, sp, Ush + t * kMaxNu, k, t, u, npost, use_pre ? zpre + (t * SW + row) * NU : nullptr);
    PROF(0);
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
    PROF(1);
    Model::template predict<ST, NM>(w, slope, stage, lane, row, x, u);
    PROF_RESET();
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      if ((a.wrap_mask >> i) & 1u) x[i] = angle_normalize<ST>(x[i]);
      if (bad) x[i] = (ST)0;
    }
    PROF(7);
    cost += (double)running_cost<ST, CT, NX, NU>(cs, x, u, part, NPARTS);
    PROF(8);
    if (a.states_out != nullptr && live && part == 0) {
      double* so = a.states_out + ((size_t)k * T + t) * NX;
#pragma unroll
      for (int i = 0; i < NX; ++i) so[i] = (double)x[i];
    }
  }
  // the mirror lanes' shares of the running cost (fixed order: part 1, 2, 3 added to part 0)
  if (NPARTS > 1) {
    const double mine = cost;
#pragma unroll
    for (int p = 1; p < NPARTS; ++p) cost += __shfl_sync(0xffffffffu, mine, (row + p * SW) & 31);
  }
  double total = cost + (double)terminal_cost<ST, CT, NX>(cs, x);
  total += (double)pert;
  if (live && part == 0) a.cost_out[k] = total;
  if (a.fuse) {
    total = __shfl_sync(0xffffffffu, total, row);  // mirrors take their sample's total from its part-0 lane
    fused_tail<NU, SW>(a, *reinterpret_cast<const SamplerS<double>*>(smem_raw + L::kSamplerD), total, row, kbase, lane, warp_index,
                       warps_total);
  }
}

// The image, the sampler and the shifted nominal sequence, staged once per CTA.
template <class Model, typename ST, typename CT>
__device__ __forceinline__ void rollout_mma_stage(const RolloutArgs& a, unsigned char* smem_raw, int tid, int nthr) {
  using L = MmaSmem<Model, ST, CT>;
  constexpr int NU = Model::NU;
  ST* Ush = reinterpret_cast<ST*>(smem_raw + L::kU);
  const int T = a.src.T;
  grid_dep_launch();  // the softmax-update launch may take its place behind this grid now
  copy_image_async(smem_raw, a.image, (int)L::kU, tid, nthr);
  for (int i = tid; i < T * NU; i += nthr) {
    int t = i / NU, d = i % NU, ts = t + a.shift;
    Ush[t * kMaxNu + d] = (ST)(ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d]);
  }
  copy_image_wait();
  __syncthreads();
}

template <class Model, typename ST, typename CT, int NM, bool PRE, int BLK = 128>
__global__ void __launch_bounds__(BLK, BLK == 256 ? 2 : 1) rollout_mma_kernel(const RolloutArgs a) {
  constexpr int SW = NM * (sizeof(typename Model::T) == 8 ? PolicyF64::MT : PolicyTF32x3::MT);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = MmaSmem<Model, ST, CT>;
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  rollout_mma_stage<Model, ST, CT>(a, smem_raw, tid, nthr);
  unsigned char* wbase = smem_raw + L::stage_off(a.src.T) + (size_t)warp * L::warp_bytes(a.src.T, SW, PRE);
  const int wi = blockIdx.x * (nthr >> 5) + warp;
  rollout_mma_warp<Model, ST, CT, NM, PRE>(a, smem_raw, wbase, wi * SW, lane, wi, gridDim.x * (nthr >> 5));
}

// One balanced wave for 56 832 < K <= 66 304 (3xTF32 / 3xFP16): the tensor pipe belongs to a sub-partition, so a wave
// lasts as long as its fullest sub-partition.  Sixteen warps per SM, the fourth warp of every sub-partition with ONE
// m-tile and the other three with two: seven 16-sample tiles per sub-partition (448 samples per SM) instead of the
// 8 | 8 | 6 | 6 that fourteen two-tile warps give.  Sample order inside the CTA: sub-partition-major.
template <class Model, typename ST, typename CT>
__global__ void __launch_bounds__(512, 1) rollout_mma_mix_kernel(const RolloutArgs a) {
  static_assert(sizeof(typename Model::T) == 4, "mixed tiles: 16-row policies");
  constexpr int MT = PolicyTF32x3::MT, PER_Q = 7 * MT, PER_CTA = 4 * PER_Q;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = MmaSmem<Model, ST, CT>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  rollout_mma_stage<Model, ST, CT>(a, smem_raw, tid, 512);
  unsigned char* wbase = smem_raw + L::stage_off(a.src.T) + (size_t)warp * L::warp_bytes(a.src.T, 2 * MT, false);
  const int q = warp & 3, j = warp >> 2;  // sub-partition, slot on it
  const int kbase = blockIdx.x * PER_CTA + q * PER_Q + j * 2 * MT;
  if (kbase >= a.K_local) return;  // (no block-wide barrier below)
  const int wi = blockIdx.x * 16 + warp, wn = gridDim.x * 16;
  if (j == 3) rollout_mma_warp<Model, ST, CT, 1, false>(a, smem_raw, wbase, kbase, lane, wi, wn);
  else rollout_mma_warp<Model, ST, CT, 2, false>(a, smem_raw, wbase, kbase, lane, wi, wn);
}

// Several balanced waves (K > one wave of the wide kernel; 3xTF32 / 3xFP16): one CTA of sixteen warps per SM, every warp
// walks the list of two-tile units (32 samples) with a stride of all warps.  The list is dealt SLOT-major — unit
// u of a pass goes to slot u / (4 SMs) of sub-partition u % (4 SMs) — so the units of the last, partial pass spread
// evenly over the sub-partitions instead of filling whole CTAs: a launch lasts ceil(units / (4 SMs)) unit times instead
// of 4 ceil(units / (16 SMs)).  K = 131 072 per GPU (the K = 1M problem sharded over 8 GPUs): 7 instead of 8 units per
// sub-partition; K = 262 144: 14 instead of 16; K = 1M: 56 either way.
template <class Model, typename ST, typename CT>
__global__ void __launch_bounds__(512, 1) rollout_mma_persist_kernel(const RolloutArgs a) {
  static_assert(sizeof(typename Model::T) == 4, "16-row policies");
  constexpr int SW = 2 * PolicyTF32x3::MT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = MmaSmem<Model, ST, CT>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  rollout_mma_stage<Model, ST, CT>(a, smem_raw, tid, 512);
  unsigned char* wbase = smem_raw + L::stage_off(a.src.T) + (size_t)warp * L::warp_bytes(a.src.T, SW, false);
  const int q = warp & 3, j = warp >> 2;  // sub-partition, slot on it
  const int lanes = gridDim.x * 4, stride = lanes * 4;
  const int units = (a.K_local + SW - 1) / SW;
#pragma unroll 1
  for (int u = j * lanes + blockIdx.x * 4 + q; u < units; u += stride) {
    rollout_mma_warp<Model, ST, CT, 2, false>(a, smem_raw, wbase, u * SW, lane, 0, 0);
    __syncwarp();  // the warp's staging rows are reused by its next unit
  }
}

// ---- two-warp pipelined variant for small K (latency-bound) ----------------------------------------------
// One CTA = one m-tile of 16 samples and two warps:
//   warp 0 ("dynamics")  guard -> network chain -> state update -> wrap, and publishes x_{t+1} into a small ring;
//   warp 1 ("cost")      trails by a step: running cost, perturbation cost, terminal cost, final write.
// The perturbed actions of the whole horizon are drawn by both warps before the loop (they do not depend on the
// state), so the horizon loop's critical path is only what the recurrence x_{t+1} = f(x_t, u_t) really needs.
template <class Model, int RING>
struct PipeSmem {
  using T = typename Model::T;
  static constexpr int SW = PolicyTF32x3::MT;
  static constexpr size_t kCost = align16<int>(Model::kBytes);
  static constexpr size_t kSampler = kCost + align16<int>(sizeof(CostS<float, float>));
  static constexpr size_t kSamplerD = kSampler + align16<int>(sizeof(SamplerS<float>));
  static constexpr size_t kU = kSamplerD + align16<int>(sizeof(SamplerS<double>));
  static __host__ __device__ constexpr size_t stage_off(int T_) { return kU + align16<int>(sizeof(float) * (size_t)T_ * kMaxNu); }
  static __host__ __device__ constexpr size_t upre_off(int T_) { return stage_off(T_) + align16<int>(sizeof(T) * Model::kStageElems); }
  static __host__ __device__ constexpr size_t ring_off(int T_) { return upre_off(T_) + align16<int>(sizeof(float) * (size_t)T_ * SW * Model::NU); }
  static __host__ __device__ constexpr size_t flags_off(int T_) { return ring_off(T_) + align16<int>(sizeof(float) * RING * SW * (Model::NX + 1)); }
  static __host__ __device__ constexpr size_t bytes(int T_) { return flags_off(T_) + 16; }
};

template <class Model>
__global__ void __launch_bounds__(64) rollout_mma_pipe_kernel(const RolloutArgs a) {
  static_assert(sizeof(typename Model::T) == 4, "pipelined variant is for the TF32x3 policy");
  constexpr int NX = Model::NX, NU = Model::NU, SW = PolicyTF32x3::MT, RING = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = PipeSmem<Model, RING>;
  unsigned char* w = smem_raw;
  CostS<float, float>& cs = *reinterpret_cast<CostS<float, float>*>(smem_raw + L::kCost);
  SamplerS<float>& sp = *reinterpret_cast<SamplerS<float>*>(smem_raw + L::kSampler);
  float* Ush = reinterpret_cast<float*>(smem_raw + L::kU);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = a.src.T;
  float* stage = reinterpret_cast<float*>(smem_raw + L::stage_off(T));
  float* upre = reinterpret_cast<float*>(smem_raw + L::upre_off(T));   // [T][SW][NU] actions the dynamics sees
  float* ring = reinterpret_cast<float*>(smem_raw + L::ring_off(T));   // [RING][SW][NX+1] x_{t+1} and the bad flag
  volatile int* flags = reinterpret_cast<volatile int*>(smem_raw + L::flags_off(T));  // [0] produced, [1] consumed
  {
    copy_image_async(smem_raw, a.image, (int)L::kU, tid, 64);
    for (int i = tid; i < T * NU; i += 64) {
      int t = i / NU, d = i % NU, ts = t + a.shift;
      Ush[t * kMaxNu + d] = (float)(ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d]);
    }
    if (tid < 2) flags[tid] = 0;
    copy_image_wait();
  }
  __syncthreads();
  const int row = lane % SW, part = lane / SW;
  const int kbase = blockIdx.x * SW;
  const bool live = kbase + row < a.K_local;
  const int k = live ? kbase + row : a.K_local - 1;
  // ---- every perturbed action of the horizon, drawn by all 64 threads (independent items)
#pragma unroll 2
  for (int item = tid; item < T * SW; item += 64) {
    const int t = item / SW, r = item % SW;
    float u[NU], npost[NU];
    sample_action<float, NU>(a.src, sp, Ush + t * kMaxNu, min(kbase + r, a.K_local - 1), t, u, npost);
#pragma unroll
    for (int i = 0; i < NU; ++i) upre[item * NU + i] = u[i];
  }
  __syncthreads();
  const float slope = (float)a.slope;
  // Warps are bound to SM sub-partitions by their index in the CTA: alternate the roles with the CTA parity so that
  // the dynamics warps (tensor-pipe heavy) of co-resident CTAs do not all share one sub-partition.
  if (((warp ^ blockIdx.x) & 1) == 0) {
    // ================= dynamics warp
    float x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = (float)a.state[i];
    const float bad2 = (float)(a.bad_norm * a.bad_norm);
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      float u[NU];
      PROF_T0();
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = upre[(t * SW + row) * NU + i];
      bool bad = false;
      if (a.has_guard) {
        float n2 = 0.f;
#pragma unroll
        for (int i = 0; i < NX; ++i) n2 = __fmaf_rn(x[i], x[i], n2);
        bad = n2 > bad2;
        if (bad) {
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = 0.f;
#pragma unroll
          for (int i = 0; i < NU; ++i) u[i] = 0.f;
        }
      }
      PROF(1);
      Model::template predict<float, 1>(w, slope, stage, lane, row, x, u);
      PROF_RESET();
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        if ((a.wrap_mask >> i) & 1u) x[i] = angle_normalize<float>(x[i]);
        if (bad) x[i] = 0.f;
      }
      PROF(7);
      if (t >= RING) {  // never ahead of the cost warp by a whole ring (it is the faster of the two)
        while (flags[1] < t - RING + 1) __nanosleep(64);
      }
      if (part == 0) {
        float* slot = ring + ((t % RING) * SW + row) * (NX + 1);
#pragma unroll
        for (int i = 0; i < NX; ++i) slot[i] = x[i];
        slot[NX] = bad ? 1.f : 0.f;
      }
      __syncwarp();
      if (lane == 0) { __threadfence_block(); flags[0] = t + 1; }
      PROF(9);
      if (a.states_out != nullptr && live && part == 0) {
        double* so = a.states_out + ((size_t)k * T + t) * NX;
#pragma unroll
        for (int i = 0; i < NX; ++i) so[i] = (double)x[i];
      }
    }
  } else {
    // ================= cost warp (two lanes per sample share the trap-set terms)
    double cost = 0.0;
    float pert = 0.f;
    float x[NX];
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      float u[NU];
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = upre[(t * SW + row) * NU + i];
      // perturbation cost of this step: sum_i u_i (lambda * (u - U_t) @ Sigma^-1)_i  (controller.py:395,400)
      if (a.src.mode == kModePhilox || a.src.mode == kModeInjected) {
        float pc = 0.f;
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          float ac = 0.f;
#pragma unroll
          for (int j = 0; j < NU; ++j) ac = __fmaf_rn(u[j] - Ush[t * kMaxNu + j], sp.lsi[j * kMaxNu + i], ac);
          pc = __fmaf_rn(u[i], ac, pc);
        }
        pert += pc;
      }
      while (flags[0] <= t) __nanosleep(64);
      __threadfence_block();
      const float* slot = ring + ((t % RING) * SW + row) * (NX + 1);
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = slot[i];
      const bool bad = slot[NX] != 0.f;
      __syncwarp();
      if (lane == 0) flags[1] = t + 1;
      if (bad) {
#pragma unroll
        for (int i = 0; i < NU; ++i) u[i] = 0.f;
      }
      cost += (double)running_cost<float, float, NX, NU>(cs, x, u, part, 2);
    }
    const double mine = cost;
    cost += __shfl_sync(0xffffffffu, mine, (row + SW) & 31);
    double total = cost + (double)terminal_cost<float, float, NX>(cs, x);
    total += (double)pert;
    if (live && part == 0) a.cost_out[k] = total;
    if (a.fuse) {
      total = __shfl_sync(0xffffffffu, total, row);
      fused_tail<NU, SW>(a, *reinterpret_cast<const SamplerS<double>*>(smem_raw + L::kSamplerD), total, row, kbase, lane, blockIdx.x,
                         gridDim.x);
    }
  }
}

// ---- role-split variant for latency-bound K (RexExtract shapes) --------------------------------------------
// At K <= ~19k there is at most one 16-sample tile per SM sub-partition and the single-warp kernel is bound by the
// length of its own dependent instruction chain (ncu: issue slots 28 % busy, stall `wait` 50 %; measured
// tools/peaks/mma_latency: HMMA.1688.TF32 issues every 8 cycles, dependent latency 20).  Here a tile is shared by two
// warps, and only what the recurrence x_{t+1} = f(x_t, u_t) really needs stays on the critical path:
//   warp A ("chain")    [x_t, u_t] -> encoder -> latent dynamics -> v_t;  x_{t+1} = x_t + C_t v_t + d0, wrap, guard
//   warp B ("decoder")  x_t -> decoder -> C_t  (ready before A's chain ends);  running / perturbation / terminal cost
// A publishes the rows of step t+1 and goes straight on; B trails: it turns the published x into C_{t+1} first and
// evaluates the cost of the finished step while A is busy.  Hand-offs: 512-byte shared-memory rows and two named
// barriers (bar.arrive / bar.sync, 64 threads).  The arithmetic per sample (layer order, accumulation order, cost
// terms) is exactly the single-warp kernel's, so both produce bitwise-identical costs.  Two tiles per 128-thread CTA;
// CTAs that land on the same SM alternate the order of the roles so that every sub-partition gets one chain warp and
// one decoder warp (warps are bound to sub-partitions by their index in the CTA).
template <class Model>
struct SplitSmem {
  using T = typename Model::T;
  static constexpr int SW = PolicyTF32x3::MT, TILES = 2;
  static constexpr int kCStride = 8 * Model::CO8 + 2;  // floats per row of the decoder output tile (+2: bank spread)
  static constexpr size_t kCost = align16<int>(Model::kBytes);
  static constexpr size_t kSampler = kCost + align16<int>(sizeof(CostS<float, float>));
  static constexpr size_t kSamplerD = kSampler + align16<int>(sizeof(SamplerS<float>));
  static constexpr size_t kU = kSamplerD + align16<int>(sizeof(SamplerS<double>));
  static __host__ __device__ constexpr size_t tiles_off(int T_) { return kU + align16<int>(sizeof(float) * (size_t)T_ * kMaxNu); }
  // per tile: xs[SW][8] network input rows | xc[SW][8] true state + bad flag | vs[SW][8] | Cs[SW][kCStride] |
  //           upre[T][SW][NU] actions | pcs[T][SW] perturbation-cost terms
  static constexpr size_t kXs = 0, kXc = sizeof(float) * SW * 8, kVs = 2 * sizeof(float) * SW * 8, kCs = 3 * sizeof(float) * SW * 8;
  static constexpr size_t kUpre = kCs + align16<int>(sizeof(float) * SW * kCStride);
  static __host__ __device__ constexpr size_t pcs_off(int T_) { return kUpre + align16<int>(sizeof(float) * (size_t)T_ * SW * Model::NU); }
  static __host__ __device__ constexpr size_t tile_bytes(int T_) { return pcs_off(T_) + align16<int>(sizeof(float) * (size_t)T_ * SW); }
  static __host__ __device__ constexpr size_t bytes(int T_) { return tiles_off(T_) + TILES * tile_bytes(T_); }
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

static __device__ unsigned int g_sm_turn[1024];  // per-SM CTA counter (never reset: only its parity is used)

// ROLES = 2: chain | decoder + cost.  ROLES = 3: chain | decoder | cost — for K small enough that every warp has a
// sub-partition to itself (<= 2 tiles per SM), where the decoder + cost warp is the longer of the two (measured at
// K=512: 64 us, 48 us without the cost); the cost warp joins both hand-offs so that nobody can run a step ahead of it.
// Unrolling of the prologue's draw loop in the role-split kernels.  Not unrolled: the bench flushes L2 between commands,
// so every launch fetches its code cold, and the prologue runs once — 1 100 fewer instructions to fetch beat the
// instruction-level parallelism of five draws in flight (K=8192: 76.1 -> 74.3 us; unroll 2: 74.6).
#ifndef TAMPC_DRAW_UNROLL
#define TAMPC_DRAW_UNROLL 1
#endif
constexpr int kDrawUnroll = TAMPC_DRAW_UNROLL;

// JIT (three roles only): the cost warp draws the actions inside the horizon loop instead of the prologue.
template <class Model, int ROLES, bool JIT = false>
__global__ void __launch_bounds__(64 * ROLES, 3) rollout_mma_split_kernel(const RolloutArgs a) {
  static_assert(!JIT || ROLES == 3, "just-in-time drawing is the cost warp's job");
  static_assert(ROLES == 2 || ROLES == 3, "two or three warps per tile");
  static_assert(sizeof(typename Model::T) == 4, "role-split variant is for the float32 tensor policies");
  static_assert(Model::NX < 8, "xc rows keep the bad flag in column NX");
  using P = typename Model::Policy;
  constexpr int NX = Model::NX, NU = Model::NU, NV = Model::kNV, SW = P::MT, CO8 = Model::CO8;
  using L = SplitSmem<Model>;
  constexpr int CS = L::kCStride;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned int s_flip;
  unsigned char* w = smem_raw;
  CostS<float, float>& cs = *reinterpret_cast<CostS<float, float>*>(smem_raw + L::kCost);
  SamplerS<float>& sp = *reinterpret_cast<SamplerS<float>*>(smem_raw + L::kSampler);
  float* Ush = reinterpret_cast<float*>(smem_raw + L::kU);
  constexpr int NTHR = 64 * ROLES, BARN = 32 * ROLES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = a.src.T;
  {
    grid_dep_launch();  // the softmax-update launch may take its place behind this grid now
    copy_image_async(smem_raw, a.image, (int)L::kU, tid, NTHR);
    if (tid == 0) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      s_flip = atomicAdd(&g_sm_turn[smid & 1023u], 1u) & 1u;
    }
    for (int i = tid; i < T * NU; i += NTHR) {
      int t = i / NU, d = i % NU, ts = t + a.shift;
      Ush[t * kMaxNu + d] = (float)(ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d]);
    }
    copy_image_wait();
  }
  __syncthreads();
  // (Three roles: reversing the (tile, role) order in every second CTA of an SM was measured and is WORSE — K=8192
  // 77 -> 97 us: six-warp CTAs already spread their chain warps over the sub-partitions.)
  const int tile = warp / ROLES, role0 = warp % ROLES;
  unsigned char* tbase = smem_raw + L::tiles_off(T) + (size_t)tile * L::tile_bytes(T);
  float* xs = reinterpret_cast<float*>(tbase + L::kXs);      // [SW][8] the step's network input rows (x_t | u_t | 0), zero if guarded
  float* xc = reinterpret_cast<float*>(tbase + L::kXc);      // [SW][8] the state the cost sees + the guard flag of the step
  float* vs = reinterpret_cast<float*>(tbase + L::kVs);      // [SW][8] latent output v_t (private to warp A)
  float* Cs = reinterpret_cast<float*>(tbase + L::kCs);      // [SW][CS] decoder output C_t
  float* upre = reinterpret_cast<float*>(tbase + L::kUpre);  // [T][SW][NU] actions the dynamics sees
  float* pcs = reinterpret_cast<float*>(tbase + L::pcs_off(T));  // [T][SW] perturbation cost of (t, sample) (controller.py:395,400)
  const int kbase = (blockIdx.x * L::TILES + tile) * SW;
  const int row = lane % SW, part = lane / SW;
  const bool live = kbase + row < a.K_local;
  const int k = live ? kbase + row : a.K_local - 1;
  // perturbed action + perturbation-cost term of item (t, r) = t * SW + r of this tile
  auto draw = [&](int item) {
    const int t = item / SW, r = item % SW;
    float u[NU], npost[NU];
    pcs[item] = sample_action<float, NU>(a.src, sp, Ush + t * kMaxNu, min(kbase + r, a.K_local - 1), t, u, npost);
#pragma unroll
    for (int i = 0; i < NU; ++i) upre[item * NU + i] = u[i];
  };
  if (kbase < a.K_local) {
    if (JIT) {
      // the cost warp has slack: it draws the actions two steps ahead inside the horizon loop (32 items = two steps
      // per draw); only steps 0 and 1 are needed before the loop starts.  Only while every warp has a sub-partition
      // to itself (<= 2 tiles per SM; measured K=500 60 -> 54 us, 4736 64 -> 57): with four tiles per SM the draws
      // take issue slots from a chain warp (K=8192 77 -> 80 us), so there they stay in the prologue (JIT = false;
      // a compile-time switch — as a run-time flag the gain was lost, 60 us at K=500).
      if (role0 == 2 && lane < T * SW) draw(lane);
    } else {
      // every perturbed action of the horizon, drawn by the tile's threads (independent items)
#pragma unroll kDrawUnroll
      for (int item = role0 * 32 + lane; item < T * SW; item += BARN) draw(item);
    }
  }
  __syncthreads();
  if (kbase >= a.K_local) return;  // (no block-wide barrier below)
  const int bar_x = 1 + 2 * tile, bar_c = 2 + 2 * tile;
  const float slope = (float)a.slope;
  // two roles: CTAs sharing an SM alternate them, so that every sub-partition gets one chain and one decoder warp
  const int role = ROLES == 2 ? (int)((role0 ^ s_flip) & 1u) : role0;

  if (role == 0) {
    // ================= warp A: encoder + latent dynamics + state update (the recurrence)
    const float bad2 = (float)(a.bad_norm * a.bad_norm);
    const float* d0 = reinterpret_cast<const float*>(w + Model::kOffD0);
    float x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = (float)a.state[i];
    bool bad = false;
    // rows of step t: guard (online_controller.py:62-66); xs = what the networks see, xc = what the cost sees
    auto publish = [&](int t) {
      bad = false;
      if (a.has_guard) {
        float n2 = 0.f;
#pragma unroll
        for (int i = 0; i < NX; ++i) n2 = __fmaf_rn(x[i], x[i], n2);
        bad = n2 > bad2;
      }
      if (part == 0) {
        float* r = xs + row * 8;
        float* c = xc + row * 8;
        if (t < T) {
          float u[NU];
#pragma unroll
          for (int i = 0; i < NU; ++i) u[i] = upre[(t * SW + row) * NU + i];
          float rv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) rv[i] = bad ? 0.f : (i < NX ? x[i] : (i < NX + NU ? u[i - NX] : 0.f));
          reinterpret_cast<float4*>(r)[0] = make_float4(rv[0], rv[1], rv[2], rv[3]);
          reinterpret_cast<float4*>(r)[1] = make_float4(rv[4], rv[5], rv[6], rv[7]);
        }
        float cv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cv[i] = i < NX ? x[i] : (i == NX && bad ? 1.f : 0.f);
        reinterpret_cast<float4*>(c)[0] = make_float4(cv[0], cv[1], cv[2], cv[3]);
        reinterpret_cast<float4*>(c)[1] = make_float4(cv[4], cv[5], cv[6], cv[7]);
      }
      __syncwarp();
      named_bar_arrive(bar_x, BARN);
    };
    publish(0);
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      PROF_T0();
      float in[1][1][P::R];
      load_blocks<P, 1, 1, 8>(xs, in, lane);
      PROF(1);
      float z[1][1][P::R], v[1][1][P::R];
      Model::Enc::template eval<1>(w, in, z, slope, lane);
      PROF(2);
      Model::Dyn::template eval<1>(w + Model::kOffDyn, z, v, slope, lane);
      PROF(3);
      store_blocks<P, 1, 1, 8>(vs, v, 0, lane);
      __syncwarp();
      named_bar_sync(bar_c, BARN);  // Cs holds C_t (and, with a cost warp, it has read the rows of step t)
      PROF(0);
      const float* cr = Cs + row * CS;
      const float* vr = vs + row * 8;
      const bool was_bad = bad;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        float dx = d0[i];
#pragma unroll
        for (int j = 0; j < NV; ++j) dx = __fmaf_rn(cr[i * NV + j], vr[j], dx);
        float xn = (was_bad ? 0.f : x[i]) + dx;
        if ((a.wrap_mask >> i) & 1u) xn = angle_normalize<float>(xn);
        x[i] = was_bad ? 0.f : xn;
      }
      __syncwarp();  // every lane is done with Cs / vs / xs of step t
      publish(t + 1);
      PROF(4);
    }
  } else if (ROLES == 3 && role == 1) {
    // ================= warp B (three roles): decoder only
#pragma unroll 1
    for (int t = 0; t <= T; ++t) {
      named_bar_sync(bar_x, BARN);
      if (t < T) {
        float in[1][1][P::R], C[1][CO8][P::R];
        load_blocks<P, 1, 1, 8>(xs, in, lane);
        Model::Dec::template eval<1>(w + Model::kOffDec, in, C, slope, lane);
        store_blocks<P, 1, CO8, CS>(Cs, C, 0, lane);
        __syncwarp();
        named_bar_arrive(bar_c, BARN);
      }
    }
  } else {
    // ================= warp B (two roles): decoder, then the cost of the step that just finished; warp C (three
    // roles): the cost alone.  Two lanes per sample share the trap-set terms.
    double cost = 0.0;
    float pert = 0.f;
    float x[NX], useen[NU];
    CostHot<float, float, NX, NU> hot;
    hot.load(cs);
#pragma unroll 1
    for (int t = 0; t <= T; ++t) {
      PROF_T0();
      named_bar_sync(bar_x, BARN);  // xs / xc hold step t
      PROF(7);
      const float* c = xc + row * 8;
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = c[i];
      const bool bad = c[NX] != 0.f;
      if (t < T) {
        if (ROLES == 2) {
          float in[1][1][P::R], C[1][CO8][P::R];
          load_blocks<P, 1, 1, 8>(xs, in, lane);
          Model::Dec::template eval<1>(w + Model::kOffDec, in, C, slope, lane);
          store_blocks<P, 1, CO8, CS>(Cs, C, 0, lane);
        }
        if (JIT && (t & 1) == 0) {  // steps t+2, t+3: read by the chain warp after the hand-offs of steps t+1, t+2
          const int item = (t + 2) * SW + lane;
          if (item < T * SW) draw(item);
        }
        __syncwarp();  // (three roles: every lane holds its row of step t before the chain warp may overwrite it)
        named_bar_arrive(bar_c, BARN);
      }
      PROF(6);
      if (t > 0) {
        // x = x_t is the state reached by step t-1 (zero if that step was guarded, online_controller.py:77)
        cost += (double)running_cost_hot<float, float, NX, NU>(cs, hot, x, useen, part, 2);
        if (a.states_out != nullptr && live && part == 0) {
          double* so = a.states_out + ((size_t)k * T + (t - 1)) * NX;
#pragma unroll
          for (int i = 0; i < NX; ++i) so[i] = (double)x[i];
        }
      }
      PROF(9);
      if (t < T) {
        pert += pcs[t * SW + row];  // drawn with the action, summed in step order
#pragma unroll
        for (int i = 0; i < NU; ++i) useen[i] = bad ? 0.f : upre[(t * SW + row) * NU + i];
      }
      PROF(5);
    }
    const double mine = cost;
    cost += __shfl_sync(0xffffffffu, mine, (row + SW) & 31);
    double total = cost + (double)terminal_cost<float, float, NX>(cs, x);
    total += (double)pert;
    if (live && part == 0) a.cost_out[k] = total;
  }
}

template <class Model, int ROLES, bool JIT>
cudaError_t launch_rollout_mma_split_as(const RolloutArgs& a, int grid, cudaStream_t stream) {
  auto kern = rollout_mma_split_kernel<Model, ROLES, JIT>;
  const size_t smem = SplitSmem<Model>::bytes(a.src.T);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  kern<<<grid, 64 * ROLES, smem, stream>>>(a);
  return cudaGetLastError();
}

template <class Model, int ROLES = 2>
cudaError_t launch_rollout_mma_split(const RolloutArgs& a, int /*block*/, cudaStream_t stream) {
  using L = SplitSmem<Model>;
  const int per_block = L::TILES * L::SW;
  const int grid = (a.K_local + per_block - 1) / per_block;
  if constexpr (ROLES == 3) {
    static const int n_sm = [] {
      int dev = 0, n = 148;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      return n;
    }();
    if (grid <= n_sm) return launch_rollout_mma_split_as<Model, 3, true>(a, grid, stream);  // at most one CTA (two tiles) per SM
  }
  return launch_rollout_mma_split_as<Model, ROLES, false>(a, grid, stream);
}

template <class Model>
cudaError_t launch_rollout_mma_pipe(const RolloutArgs& a, int /*block*/, cudaStream_t stream) {
  auto kern = rollout_mma_pipe_kernel<Model>;
  const size_t smem = PipeSmem<Model, 4>::bytes(a.src.T);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int grid = (a.K_local + PolicyTF32x3::MT - 1) / PolicyTF32x3::MT;
  kern<<<grid, 64, smem, stream>>>(a);
  return cudaGetLastError();
}

template <class Model, typename ST, typename CT, int NM, bool PRE, int BLK = 128>
cudaError_t launch_rollout_mma(const RolloutArgs& a, int block, cudaStream_t stream) {
  auto kern = rollout_mma_kernel<Model, ST, CT, NM, PRE, BLK>;
  if (block > BLK) block = BLK;
  constexpr int SW = NM * (sizeof(typename Model::T) == 8 ? PolicyF64::MT : PolicyTF32x3::MT);
  const int warps = block / 32;
  const size_t smem = MmaSmem<Model, ST, CT>::bytes(a.src.T, warps, SW, PRE);
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

template <class Model, typename ST, typename CT>
cudaError_t launch_rollout_mma_mix(const RolloutArgs& a, int /*block*/, cudaStream_t stream) {
  auto kern = rollout_mma_mix_kernel<Model, ST, CT>;
  if (a.fuse) return cudaErrorInvalidValue;  // the fused tail numbers warps by equal sample counts
  const size_t smem = MmaSmem<Model, ST, CT>::bytes(a.src.T, 16, 2 * PolicyTF32x3::MT, false);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  constexpr int per_block = 4 * 7 * PolicyTF32x3::MT;
  kern<<<(a.K_local + per_block - 1) / per_block, 512, smem, stream>>>(a);
  return cudaGetLastError();
}

template <class Model, typename ST, typename CT>
cudaError_t launch_rollout_mma_persist(const RolloutArgs& a, int /*block*/, cudaStream_t stream) {
  auto kern = rollout_mma_persist_kernel<Model, ST, CT>;
  if (a.fuse) return cudaErrorInvalidValue;
  const size_t smem = MmaSmem<Model, ST, CT>::bytes(a.src.T, 16, 2 * PolicyTF32x3::MT, false);
  static size_t configured = 0;
  static int n_sm = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  if (n_sm == 0) {
    int dev = 0;
    n_sm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  kern<<<n_sm, 512, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace tampc
