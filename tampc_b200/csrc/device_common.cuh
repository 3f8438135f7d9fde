// device_common.cuh — device-side building blocks of the fused MPPI rollout (sm_100a).
//
// What the reference does per horizon step with ~60-100 ATen launches on (M*K, .) float64 tensors
// (tampc/controller/controller.py:359-368) is done here per sample entirely in registers:
//   sample_action()   controller.py:386-395  noise draw, perturb, clamp, post-clamp noise, perturbation cost
//   guard + predict   online_controller.py:60-79, dynamics/model.py:147-170, transform/invariant.py:792-795,926-931
//   running_cost()    cost.py:11-36,154-189,213-235 (+ recovery cost.py:114-147)
// Weights are staged once per CTA in shared memory and read with broadcast LDS.128; float32 networks use the
// packed `fma.rn.f32x2` (FFMA2) pipe with output neurons paired in 64-bit registers (measured on B200,
// tools/peaks: 56.9 TFLOP/s = 77% of the 74 TFLOP/s FP32 peak for a 32x32 layer at 2 samples/thread).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/tampc_mppi.h"

namespace tampc {

// Programmatic dependent launch (sm_90+): the softmax-update launch that follows a rollout is placed on the SMs while
// the rollout still runs and blocks in grid_dep_wait() until the rollout grid has completed and its stores are visible.
// Both are no-ops in a launch without the programmatic-serialization attribute.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Flag-in-data publication of a double to memory somebody else polls (mapped pinned host memory; peer memory): two
// 8-byte words, each = 32 bits of the value | the 32-bit sequence number of the command.  An aligned 8-byte store is
// never torn, so the reader needs no flag behind the data and the writer no fence in front of a flag: it polls the
// words themselves until both carry the sequence number it waits for.  slot = index of the value (16 bytes apart).
constexpr int kHostSlots = 8;  // action[kMaxNu], beta, eta, argmin, error
__device__ __forceinline__ void ll_publish_bits(unsigned long long* base, int slot, unsigned long long b, unsigned long long seq) {
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(base) + 2 * slot;
  w[0] = (b & 0xffffffffull) | tag;
  w[1] = (b >> 32) | tag;
}
__device__ __forceinline__ void ll_publish(unsigned long long* base, int slot, double v, unsigned long long seq) {
  ll_publish_bits(base, slot, (unsigned long long)__double_as_longlong(v), seq);
}
// the value of a slot once both of its words carry seq
__device__ __forceinline__ bool ll_try_read(const unsigned long long* base, int slot, unsigned long long seq, unsigned long long* bits) {
  const volatile unsigned long long* w = reinterpret_cast<const volatile unsigned long long*>(base) + 2 * slot;
  const unsigned long long w0 = w[0], w1 = w[1], tag = seq & 0xffffffffull;
  if ((w0 >> 32) != tag || (w1 >> 32) != tag) return false;
  *bits = (w0 & 0xffffffffull) | (w1 << 32);
  return true;
}

constexpr int kMaxNx = TAMPC_MAX_NX;
constexpr int kMaxNu = TAMPC_MAX_NU;
constexpr double kPi = 3.141592653589793;
constexpr double kTrapMaxCost = 100000.0;  // cost.py:151
constexpr double kCosSimEps = 1e-8;        // torch.cosine_similarity eps

enum SampleMode : int { kModePhilox = 0, kModeInjected = 1, kModeActions = 2, kModeNominal = 3 };

// ---------------------------------------------------------------------------------------------
// sampler constants (device master copy in float64; converted to the state type when staged)
struct SamplerD {
  double mu[kMaxNu];
  double chol[kMaxNu * kMaxNu];           // lower-triangular L, Sigma = L L^T (row-major nu*nu packed in 4x4)
  double lam_sigma_inv[kMaxNu * kMaxNu];  // lambda * Sigma^-1 (row-major, 4x4)
  double u_min[kMaxNu], u_max[kMaxNu], u_init[kMaxNu];
  double lambda_inv;
  int has_bounds, null_action;
};

template <typename T>
struct SamplerS {
  T mu[kMaxNu], chol[kMaxNu * kMaxNu], lsi[kMaxNu * kMaxNu], u_min[kMaxNu], u_max[kMaxNu];
  int has_bounds, null_action;
};

// Cost program staged in shared memory.  ST = state type (positions are differenced in the state's precision),
// CT = arithmetic type of the cost terms (float for the float32-network modes, double for F64).
template <typename ST, typename CT>
struct CostS {
  int kind, use_trap, n_traps, has_terminal, n_sets;
  int set_size[TAMPC_MAX_GOAL_SETS];
  uint32_t ang_mask, usim_mask;
  ST goal[kMaxNx];
  ST trap_x[TAMPC_MAX_TRAPS][kMaxNx];
  ST set_goals[TAMPC_MAX_GOAL_SETS][TAMPC_MAX_SET_GOALS][kMaxNx];
  CT dist_scale[kMaxNx];
  CT Q[kMaxNx * kMaxNx], R[kMaxNu * kMaxNu];
  CT trap_u[TAMPC_MAX_TRAPS][kMaxNu];
  CT trap_u_norm[TAMPC_MAX_TRAPS];  // max(||trap_u[usim dims]||, eps), precomputed once per CTA
  CT trap_weight, terminal_multiplier;
  CT set_weight[TAMPC_MAX_GOAL_SETS];
  CT recovery_scale;
  // APF program (cost.py:236-289)
  int apf_use_goal, apf_helicoid, apf_clockwise;
  CT apf_max_dist, apf_inv_max_dist, apf_gain, apf_orig_angle;
  ST apf_oxy[2];
};

// ---------------------------------------------------------------------------------------------
// math helpers, overloaded on float / double
__device__ __forceinline__ float t_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double t_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float t_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double t_sqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float t_fmod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double t_fmod(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ float t_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double t_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float t_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double t_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float t_sin(float a) { return sinf(a); }
__device__ __forceinline__ double t_sin(double a) { return sin(a); }

// cost-term division and square root.  float: MUFU-based (<= 2 ulp); double: IEEE
__device__ __forceinline__ float t_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double t_div(double a, double b) { return a / b; }
__device__ __forceinline__ float t_sqrt_fast(float a) { return a * rsqrtf(fmaxf(a, 1e-30f)); }
__device__ __forceinline__ double t_sqrt_fast(double a) { return sqrt(a); }
__device__ __forceinline__ float t_floor(float a) { return floorf(a); }
__device__ __forceinline__ double t_floor(double a) { return floor(a); }
__device__ __forceinline__ float t_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double t_atan2(double y, double x) { return atan2(y, x); }

// torch.remainder(a + pi, 2pi) - pi  (math_utils.angle_normalize; scripts/push_main.py:192), computed as
// y - 2pi*floor(y/2pi) with one correction step instead of the (slow, looping) fmod; agrees with the fmod form
// to an ulp of 2pi.
template <typename T>
__device__ __forceinline__ T angle_normalize(T a) {
  // branch-free; an input already in [-pi, pi) takes k = 0 and comes back as (a + pi) - pi, the reference's rounding
  const T two_pi = (T)(2.0 * kPi);
  const T y = a + (T)kPi;
  T r = t_fma(-t_floor(y * (T)(1.0 / (2.0 * kPi))), two_pi, y);
  r = r < (T)0 ? r + two_pi : (r >= two_pi ? r - two_pi : r);
  return r - (T)kPi;
}

// math_utils.angular_diff_batch: one wrap into [-pi, pi] (tampc/env/block_push.py:957)
template <typename T>
__device__ __forceinline__ T angular_wrap_once(T d) {
  const T shift = d > (T)kPi ? (T)(-2.0 * kPi) : (d < (T)(-kPi) ? (T)(2.0 * kPi) : (T)0);  // selects, no branches
  return d + shift;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (global sample, t, call lo, call hi), key = seed
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// four standard normals from one Philox block (Box-Muller in float32; u in (0,1))
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t call, uint32_t sample, uint32_t t, float (&z)[4]) {
  uint4 r = philox4x32_10(make_uint4(sample, t, (uint32_t)call, (uint32_t)(call >> 32)),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float s = 2.3283064365386963e-10f;  // 2^-32
  float u0 = __fmaf_rn((float)r.x, s, 0.5f * s), u1 = __fmaf_rn((float)r.y, s, 0.5f * s);
  float u2 = __fmaf_rn((float)r.z, s, 0.5f * s), u3 = __fmaf_rn((float)r.w, s, 0.5f * s);
  u0 = fminf(u0, 0.99999994f);
  u2 = fminf(u2, 0.99999994f);
  // fast intrinsics (MUFU): |error| ~ 2^-21, far below anything a Gaussian perturbation can resolve; the phase is
  // taken in [-pi, pi) where __sincosf is most accurate (a uniform phase shifted by pi is still uniform)
  float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.2831853071795865f * u1 - 3.14159265358979324f, &s0, &c0);
  __sincosf(6.2831853071795865f * u3 - 3.14159265358979324f, &s1, &c1);
  z[0] = r0 * c0;
  z[1] = r0 * s0;
  z[2] = r1 * c1;
  z[3] = r1 * s1;
}

// ---------------------------------------------------------------------------------------------
// sampling of the perturbed action for (sample k, step t)      [controller.py:386-395]
struct SampleSrc {
  int mode;
  const double* noise;    // kModeInjected: (K_local, T, nu) draw N(mu, Sigma)
  const double* actions;  // kModeActions:  (K_local, T, nu) final actions
  uint64_t seed, call;
  int sample_offset, k_global, T;
};

// u[]: the action the dynamics sees; npost[]: post-clamp noise (perturbed - U[t]); returns the perturbation-cost
// increment  sum_i perturbed_i * (lambda * npost @ Sigma^-1)_i   (controller.py:395,400)
template <typename T, int NU>
__device__ __forceinline__ T sample_action(const SampleSrc& src, const SamplerS<T>& sp, const T* __restrict__ Ut,
                                           int k_local, int t, T (&u)[NU], T (&npost)[NU], const float* __restrict__ zpre = nullptr) {
  if (src.mode == kModeNominal) {
#pragma unroll
    for (int i = 0; i < NU; ++i) { u[i] = Ut[i]; npost[i] = (T)0; }
    return (T)0;
  }
  if (src.mode == kModeActions) {
    const double* a = src.actions + ((size_t)k_local * src.T + t) * NU;
#pragma unroll
    for (int i = 0; i < NU; ++i) { u[i] = (T)a[i]; npost[i] = (T)0; }
    return (T)0;
  }
  T n[NU];
  if (src.mode == kModeInjected) {
    const double* p = src.noise + ((size_t)k_local * src.T + t) * NU;
#pragma unroll
    for (int i = 0; i < NU; ++i) n[i] = (T)p[i];
  } else {
    float z[4];
    if (zpre != nullptr) {  // standard normals drawn ahead of the horizon loop (same Philox stream)
#pragma unroll
      for (int i = 0; i < NU; ++i) z[i] = zpre[i];
    } else {
      philox_normal4(src.seed, src.call, (uint32_t)(src.sample_offset + k_local), (uint32_t)t, z);
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      T acc = sp.mu[i];
#pragma unroll
      for (int j = 0; j <= i; ++j) acc = t_fma(sp.chol[i * kMaxNu + j], (T)z[j], acc);
      n[i] = acc;
    }
  }
  const bool null_row = sp.null_action && (src.sample_offset + k_local == src.k_global - 1);
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    T p = null_row ? (T)0 : (Ut[i] + n[i]);
    if (sp.has_bounds) p = t_max(t_min(p, sp.u_max[i]), sp.u_min[i]);
    u[i] = p;
    npost[i] = p - Ut[i];
  }
  T pc = (T)0;
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    T ac = (T)0;
#pragma unroll
    for (int j = 0; j < NU; ++j) ac = t_fma(npost[j], sp.lsi[j * kMaxNu + i], ac);
    pc = t_fma(u[i], ac, pc);
  }
  return pc;
}

// ---------------------------------------------------------------------------------------------
// packed float32 helpers
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// Padded output width of a packed layer block: float rows are multiples of 4 (one LDS.128), double rows of 2.
template <typename NT> __host__ __device__ constexpr int pad_out(int out) { return sizeof(NT) == 4 ? ((out + 3) & ~3) : ((out + 1) & ~1); }
// Elements of one layer block: row 0 = bias, row i+1 = weights of input i to every (padded) output.
template <typename NT> __host__ __device__ constexpr int layer_elems(int in, int out) { return (in + 1) * pad_out<NT>(out); }

// y = act(W x + b) for S samples held by this thread; weights broadcast from shared memory.
template <int S, int IN, int OUT, bool ACT>
__device__ __forceinline__ void linear(const float (&x)[S][IN], float (&y)[S][OUT], const float* __restrict__ w, float slope) {
  constexpr int OP = pad_out<float>(OUT);
  const ulonglong2* __restrict__ w2 = reinterpret_cast<const ulonglong2*>(w);
  uint64_t acc[S][OP / 2];
#pragma unroll
  for (int j4 = 0; j4 < OP / 4; ++j4) {
    ulonglong2 b = w2[j4];
#pragma unroll
    for (int s = 0; s < S; ++s) { acc[s][2 * j4] = b.x; acc[s][2 * j4 + 1] = b.y; }
  }
#pragma unroll
  for (int i = 0; i < IN; ++i) {
    uint64_t xx[S];
#pragma unroll
    for (int s = 0; s < S; ++s) xx[s] = pack2(x[s][i], x[s][i]);
#pragma unroll
    for (int j4 = 0; j4 < OP / 4; ++j4) {
      ulonglong2 ww = w2[(i + 1) * (OP / 4) + j4];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        acc[s][2 * j4] = ffma2(xx[s], ww.x, acc[s][2 * j4]);
        acc[s][2 * j4 + 1] = ffma2(xx[s], ww.y, acc[s][2 * j4 + 1]);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j2 = 0; j2 < OP / 2; ++j2) {
      float a, b;
      unpack2(acc[s][j2], a, b);
      if (ACT) { a = fmaxf(a, slope * a); b = fmaxf(b, slope * b); }
      if (2 * j2 < OUT) y[s][2 * j2] = a;
      if (2 * j2 + 1 < OUT) y[s][2 * j2 + 1] = b;
    }
}

template <int S, int IN, int OUT, bool ACT>
__device__ __forceinline__ void linear(const double (&x)[S][IN], double (&y)[S][OUT], const double* __restrict__ w, double slope) {
  constexpr int OP = pad_out<double>(OUT);
  const double2* __restrict__ w2 = reinterpret_cast<const double2*>(w);
  double acc[S][OP];
#pragma unroll
  for (int j2 = 0; j2 < OP / 2; ++j2) {
    double2 b = w2[j2];
#pragma unroll
    for (int s = 0; s < S; ++s) { acc[s][2 * j2] = b.x; acc[s][2 * j2 + 1] = b.y; }
  }
#pragma unroll
  for (int i = 0; i < IN; ++i) {
#pragma unroll
    for (int j2 = 0; j2 < OP / 2; ++j2) {
      double2 ww = w2[(i + 1) * (OP / 2) + j2];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        acc[s][2 * j2] = __fma_rn(x[s][i], ww.x, acc[s][2 * j2]);
        acc[s][2 * j2 + 1] = __fma_rn(x[s][i], ww.y, acc[s][2 * j2 + 1]);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int j = 0; j < OUT; ++j) {
      double a = acc[s][j];
      if (ACT) a = fmax(a, slope * a);
      y[s][j] = a;
    }
}

// Three-layer MLP  I -> A -> B -> O  (every network on the reference path has this shape)
template <typename NT, int I, int A, int B, int O>
struct Mlp3 {
  static constexpr int kIn = I, kOut = O;
  static constexpr int kOff1 = layer_elems<NT>(I, A);
  static constexpr int kOff2 = kOff1 + layer_elems<NT>(A, B);
  static constexpr int kElems = kOff2 + layer_elems<NT>(B, O);
  static constexpr int kMacs = I * A + A * B + B * O;
  template <int S>
  static __device__ __forceinline__ void eval(const NT* __restrict__ w, const NT (&x)[S][I], NT (&y)[S][O], NT slope) {
    NT h1[S][A];
    linear<S, I, A, true>(x, h1, w, slope);
    NT h2[S][B];
    linear<S, A, B, true>(h1, h2, w + kOff1, slope);
    linear<S, B, O, false>(h2, y, w + kOff2, slope);
  }
};

// ---------------------------------------------------------------------------------------------
// dynamics models: predict() advances S samples one step (no guard / wrap; those are applied by the caller)
//
// ModelMlp: x' = x + net'([x,u]) with the input/output scalers folded into the first/last layer on the host.
template <typename NT, int NX_, int NU_, int H1, int H2>
struct ModelMlp {
  using Net = Mlp3<NT, NX_ + NU_, H1, H2, NX_>;
  static constexpr int NX = NX_, NU = NU_;
  static constexpr int kElems = Net::kElems;
  static constexpr int kMacs = Net::kMacs;
  template <typename ST, int S>
  static __device__ __forceinline__ void predict(const NT* __restrict__ w, NT slope, ST (&x)[S][NX], const ST (&u)[S][NU]) {
    NT in[S][NX + NU];
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int i = 0; i < NX; ++i) in[s][i] = (NT)x[s][i];
#pragma unroll
      for (int i = 0; i < NU; ++i) in[s][NX + i] = (NT)u[s][i];
    }
    NT dx[S][NX];
    Net::template eval<S>(w, in, dx, slope);
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int i = 0; i < NX; ++i) x[s][i] += (ST)dx[s][i];
  }
};

// ModelRex: RexExtract (transform/invariant.py:792-795,926-931) + latent dynamics (dynamics/model.py:300-301).
// Host-side folding (mppi_api.cu pack_rex): z-scaler into the encoder's last layer, v un-scaler into the latent
// net's last layer, x_extractor into the decoder's first layer, the 1/y-scale into the decoder's last layer;
// what is left is three MLPs, dx = C v + d0 and x' = x + dx.
template <typename NT, int NX_, int NU_, int E1, int E2, int NZ, int D1, int D2, int NV, int C1, int C2>
struct ModelRex {
  using Enc = Mlp3<NT, NX_ + NU_, E1, E2, NZ>;
  using Dyn = Mlp3<NT, NZ, D1, D2, NV>;
  using Dec = Mlp3<NT, NX_, C1, C2, NX_ * NV>;
  static constexpr int NX = NX_, NU = NU_;
  static constexpr int kOffDyn = Enc::kElems;
  static constexpr int kOffDec = kOffDyn + Dyn::kElems;
  static constexpr int kOffD0 = kOffDec + Dec::kElems;
  static constexpr int kElems = kOffD0 + ((NX_ + 3) & ~3);
  static constexpr int kMacs = Enc::kMacs + Dyn::kMacs + Dec::kMacs + NX_ * NV;
  template <typename ST, int S>
  static __device__ __forceinline__ void predict(const NT* __restrict__ w, NT slope, ST (&x)[S][NX], const ST (&u)[S][NU]) {
    NT v[S][NV];
    {
      NT in[S][NX + NU];
#pragma unroll
      for (int s = 0; s < S; ++s) {
#pragma unroll
        for (int i = 0; i < NX; ++i) in[s][i] = (NT)x[s][i];
#pragma unroll
        for (int i = 0; i < NU; ++i) in[s][NX + i] = (NT)u[s][i];
      }
      NT z[S][NZ];
      Enc::template eval<S>(w, in, z, slope);
      Dyn::template eval<S>(w + kOffDyn, z, v, slope);
    }
    NT xin[S][NX];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int i = 0; i < NX; ++i) xin[s][i] = (NT)x[s][i];
    NT C[S][NX * NV];
    Dec::template eval<S>(w + kOffDec, xin, C, slope);
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        NT dx = w[kOffD0 + i];
#pragma unroll
        for (int j = 0; j < NV; ++j) dx = t_fma(C[s][i * NV + j], v[s][j], dx);
        x[s][i] += (ST)dx;
      }
  }
};

// ModelPendulum: scripts/pendulum.py:223-241 (params: g, m, l, dt, |u| clamp, |thdot| clamp in w[0..5])
template <typename NT>
struct ModelPendulum {
  static constexpr int NX = 2, NU = 1;
  static constexpr int kElems = 8;
  static constexpr int kMacs = 10;
  template <typename ST, int S>
  static __device__ __forceinline__ void predict(const NT* __restrict__ w, NT, ST (&x)[S][2], const ST (&u)[S][1]) {
    const ST g = (ST)w[0], m = (ST)w[1], l = (ST)w[2], dt = (ST)w[3], umax = (ST)w[4], thdmax = (ST)w[5];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      ST th = x[s][0], thdot = x[s][1];
      ST uc = t_max(t_min(u[s][0], umax), -umax);
      // newthdot = thdot + (-3g/(2l) * sin(th + pi) + 3/(m l^2) * u) * dt   (same operation order as torch)
      ST a = ((ST)-3 * g / ((ST)2 * l)) * t_sin(th + (ST)kPi) + ((ST)3 / (m * l * l)) * uc;
      ST newthdot = thdot + a * dt;
      ST newth = th + newthdot * dt;
      newthdot = t_max(t_min(newthdot, thdmax), -thdmax);
      x[s][0] = newth;
      x[s][1] = newthdot;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// cost
// env.state_difference: x - goal in the state's precision, single-wrap angular dims, then handed to the cost type
template <typename ST, typename CT, int NX>
__device__ __forceinline__ void state_diff(uint32_t ang_mask, const ST (&x)[NX], const ST* __restrict__ goal, CT (&d)[NX]) {
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    ST v = x[i] - goal[i];
    if ((ang_mask >> i) & 1u) v = angular_wrap_once(v);
    d[i] = (CT)v;
  }
}

// squared state distance  ||diff * dist_scale||^2.  The reference takes the norm and squares it again
// (cost.py:135,175); the sqrt/square round trip (<= 2 ulp) is not reproduced.
template <typename CT, int NX>
__device__ __forceinline__ CT dist_sq(const CT* __restrict__ dist_scale, const CT (&d)[NX]) {
  CT s = (CT)0;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    CT v = d[i] * dist_scale[i];
    s = t_fma(v, v, s);
  }
  return s;
}

template <typename CT, int NX>
__device__ __forceinline__ CT quad_form_x(const CT* __restrict__ Q, const CT (&d)[NX]) {
  CT total = (CT)0;
#pragma unroll
  for (int j = 0; j < NX; ++j) {
    CT col = (CT)0;  // (d @ Q)_j
#pragma unroll
    for (int i = 0; i < NX; ++i) col = t_fma(d[i], Q[i * kMaxNx + j], col);
    total = t_fma(col, d[j], total);
  }
  return total;
}

// running cost of (x_next, u)      [controller.py:248-249 -> cost.py:213-235]
// `part` of `nparts`: lanes that mirror the same sample share the work — the trap-set terms are dealt round-robin,
// everything else goes to part 0; the caller adds the parts' sums (once, after the horizon loop).
// APFVO / APFSP potentials of a (next) state (online_controller.py:564,586,639-662; cost.py:236-289): optional goal cost
// d'Qd (U=None: no action term), artificial repulsion from the trap states, optional helicoid around the active obstacle.
template <typename ST, typename CT, int NX>
__device__ __forceinline__ CT apf_cost(const CostS<ST, CT>& c, const ST (&x)[NX]) {
  CT d[NX];
  CT cost = (CT)0;
  if (c.apf_use_goal) {
    state_diff<ST, CT, NX>(c.ang_mask, x, c.goal, d);
    cost = quad_form_x<CT, NX>(c.Q, d);
  }
  CT rep = (CT)0;
  for (int p = 0; p < c.n_traps; ++p) {
    state_diff<ST, CT, NX>(c.ang_mask, x, c.trap_x[p], d);
    const CT rho = t_sqrt(dist_sq<CT, NX>(c.dist_scale, d));
    CT v = (CT)1 / rho - c.apf_inv_max_dist;   // rho = 0 -> inf -> clamped, like torch
    v = t_max(t_min(v * v, (CT)kTrapMaxCost), (CT)0);
    rep += rho <= c.apf_max_dist ? v : (CT)0;
  }
  cost = t_fma(rep, c.apf_gain, cost);
  if (c.apf_helicoid) {
    const CT dx = (CT)(c.apf_oxy[0] - x[0]), dy = (CT)(c.apf_oxy[1] - x[1]);
    const CT ang = angular_wrap_once<CT>(c.apf_orig_angle - t_atan2(dy, dx));
    const CT hel = ang * t_sqrt(t_fma(dx, dx, dy * dy));
    cost += c.apf_clockwise ? hel : -hel;
  }
  return cost;
}

template <typename ST, typename CT, int NX, int NU>
__device__ __forceinline__ CT running_cost(const CostS<ST, CT>& c, const ST (&x)[NX], const ST (&us)[NU], int part = 0, int nparts = 1) {
  CT d[NX];
  if (c.kind == TAMPC_COST_APF) return part == 0 ? apf_cost<ST, CT, NX>(c, x) : (CT)0;
  if (c.kind == TAMPC_COST_RECOVERY) {
    if (part != 0) return (CT)0;
    CT total = (CT)0;
    for (int j = 0; j < c.n_sets; ++j) {
      CT best = (CT)0;
      for (int g = 0; g < c.set_size[j]; ++g) {
        state_diff<ST, CT, NX>(c.ang_mask, x, c.set_goals[j][g], d);
        CT v = dist_sq<CT, NX>(c.dist_scale, d) * c.recovery_scale;
        best = (g == 0) ? v : t_min(best, v);
      }
      total += best * c.set_weight[j];
    }
    return total;
  }
  CT u[NU];
#pragma unroll
  for (int i = 0; i < NU; ++i) u[i] = (CT)us[i];
  CT cost = (CT)0;
  if (part == 0) {
    state_diff<ST, CT, NX>(c.ang_mask, x, c.goal, d);
    cost = quad_form_x<CT, NX>(c.Q, d);
    CT ru = (CT)0;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      CT col = (CT)0;
#pragma unroll
      for (int i = 0; i < NU; ++i) col = t_fma(u[i], c.R[i * kMaxNu + j], col);
      ru = t_fma(col, u[j], ru);
    }
    cost += ru;
  }
  if (c.use_trap && c.n_traps > 0) {
    // ||u[usim dims]|| once per step
    CT un = (CT)0;
#pragma unroll
    for (int i = 0; i < NU; ++i)
      if ((c.usim_mask >> i) & 1u) un = t_fma(u[i], u[i], un);
    un = t_max(t_sqrt_fast(un), (CT)kCosSimEps);
    CT tsum = (CT)0;
    for (int p = part; p < c.n_traps; p += nparts) {
      state_diff<ST, CT, NX>(c.ang_mask, x, c.trap_x[p], d);
      CT cx = t_max(t_min(t_div((CT)1, dist_sq<CT, NX>(c.dist_scale, d)), (CT)kTrapMaxCost), (CT)0);
      CT dot = (CT)0;
#pragma unroll
      for (int i = 0; i < NU; ++i)
        if ((c.usim_mask >> i) & 1u) dot = t_fma(u[i], c.trap_u[p][i], dot);
      CT cu = t_max(t_min(t_div(dot, un * c.trap_u_norm[p]), (CT)1), (CT)0);
      tsum = t_fma(cx, cu, tsum);
    }
    cost = t_fma(tsum, c.trap_weight, cost);
  }
  return cost;
}

// The loop-invariant operands of the goal / trap cost held in registers.  A rollout loop that stores to shared
// memory between steps cannot keep them there by itself (the compiler must assume the stores alias the cost block),
// so every step would re-load ~45 scalars; latency-bound kernels load them once before the horizon loop instead.
template <typename ST, typename CT, int NX, int NU>
struct CostHot {
  ST goal[NX];
  CT Q[NX * NX], R[NU * NU], dist_scale[NX];
  __device__ __forceinline__ void load(const CostS<ST, CT>& c) {
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      goal[i] = c.goal[i];
      dist_scale[i] = c.dist_scale[i];
#pragma unroll
      for (int j = 0; j < NX; ++j) Q[i * NX + j] = c.Q[i * kMaxNx + j];
    }
#pragma unroll
    for (int i = 0; i < NU; ++i)
#pragma unroll
      for (int j = 0; j < NU; ++j) R[i * NU + j] = c.R[i * kMaxNu + j];
  }
};

// running_cost with the goal-cost operands taken from registers; same operations in the same order
template <typename ST, typename CT, int NX, int NU>
__device__ __forceinline__ CT running_cost_hot(const CostS<ST, CT>& c, const CostHot<ST, CT, NX, NU>& h, const ST (&x)[NX], const ST (&us)[NU],
                                               int part, int nparts) {
  if (c.kind != TAMPC_COST_GOAL_TRAP) return running_cost<ST, CT, NX, NU>(c, x, us, part, nparts);
  CT d[NX];
  CT u[NU];
#pragma unroll
  for (int i = 0; i < NU; ++i) u[i] = (CT)us[i];
  CT cost = (CT)0;
  if (part == 0) {
    state_diff<ST, CT, NX>(c.ang_mask, x, h.goal, d);
    CT total = (CT)0;
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      CT col = (CT)0;
#pragma unroll
      for (int i = 0; i < NX; ++i) col = t_fma(d[i], h.Q[i * NX + j], col);
      total = t_fma(col, d[j], total);
    }
    cost = total;
    CT ru = (CT)0;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
      CT col = (CT)0;
#pragma unroll
      for (int i = 0; i < NU; ++i) col = t_fma(u[i], h.R[i * NU + j], col);
      ru = t_fma(col, u[j], ru);
    }
    cost += ru;
  }
  if (c.use_trap && c.n_traps > 0) {
    CT un = (CT)0;
#pragma unroll
    for (int i = 0; i < NU; ++i)
      if ((c.usim_mask >> i) & 1u) un = t_fma(u[i], u[i], un);
    un = t_max(t_sqrt_fast(un), (CT)kCosSimEps);
    CT tsum = (CT)0;
    for (int p = part; p < c.n_traps; p += nparts) {
      state_diff<ST, CT, NX>(c.ang_mask, x, c.trap_x[p], d);
      CT cx = t_max(t_min(t_div((CT)1, dist_sq<CT, NX>(h.dist_scale, d)), (CT)kTrapMaxCost), (CT)0);
      CT dot = (CT)0;
#pragma unroll
      for (int i = 0; i < NU; ++i)
        if ((c.usim_mask >> i) & 1u) dot = t_fma(u[i], c.trap_u[p][i], dot);
      CT cu = t_max(t_min(t_div(dot, un * c.trap_u_norm[p]), (CT)1), (CT)0);
      tsum = t_fma(cx, cu, tsum);
    }
    cost = t_fma(tsum, c.trap_weight, cost);
  }
  return cost;
}

// terminal_cost_multiplier * cost(x_T, terminal=True): QR on the state only; the trap term is multiplied by
// c_u = 0 when terminal (cost.py:177)      [controller.py:251-259]
template <typename ST, typename CT, int NX>
__device__ __forceinline__ CT terminal_cost(const CostS<ST, CT>& c, const ST (&x)[NX]) {
  if (c.kind != TAMPC_COST_GOAL_TRAP || !c.has_terminal) return (CT)0;
  CT d[NX];
  state_diff<ST, CT, NX>(c.ang_mask, x, c.goal, d);
  return c.terminal_multiplier * quad_form_x<CT, NX>(c.Q, d);
}

// ---------------------------------------------------------------------------------------------
// staging of the small constant blocks into shared memory (float64 master -> T)
template <typename T>
__device__ __forceinline__ void stage_sampler(SamplerS<T>& dst, const SamplerD& src, int tid, int nthreads) {
  for (int i = tid; i < kMaxNu; i += nthreads) {
    dst.mu[i] = (T)src.mu[i];
    dst.u_min[i] = (T)src.u_min[i];
    dst.u_max[i] = (T)src.u_max[i];
  }
  for (int i = tid; i < kMaxNu * kMaxNu; i += nthreads) {
    dst.chol[i] = (T)src.chol[i];
    dst.lsi[i] = (T)src.lam_sigma_inv[i];
  }
  if (tid == 0) { dst.has_bounds = src.has_bounds; dst.null_action = src.null_action; }
}

template <typename ST, typename CT>
__device__ __forceinline__ void stage_cost(CostS<ST, CT>& dst, const tampc_cost_desc& src, int tid, int nthreads) {
  if (tid == 0) {
    dst.kind = src.kind; dst.use_trap = src.use_trap_cost; dst.n_traps = src.n_traps;
    dst.has_terminal = src.has_terminal; dst.n_sets = src.n_sets;
    for (int j = 0; j < TAMPC_MAX_GOAL_SETS; ++j) { dst.set_size[j] = src.set_size[j]; dst.set_weight[j] = (CT)src.set_weight[j]; }
    dst.ang_mask = src.ang_mask; dst.usim_mask = src.usim_mask;
    dst.trap_weight = (CT)src.trap_weight; dst.terminal_multiplier = (CT)src.terminal_multiplier;
    dst.recovery_scale = (CT)src.recovery_scale;
    dst.apf_use_goal = src.apf_use_goal; dst.apf_helicoid = src.apf_helicoid; dst.apf_clockwise = src.apf_clockwise;
    dst.apf_max_dist = (CT)src.apf_max_dist; dst.apf_inv_max_dist = (CT)(1.0 / src.apf_max_dist); dst.apf_gain = (CT)src.apf_gain;
    // original_angle = atan2 of (oxy - rxy) (cost.py:277-278), once per program
    dst.apf_orig_angle = (CT)atan2(src.apf_oxy[1] - src.apf_rxy[1], src.apf_oxy[0] - src.apf_rxy[0]);
    dst.apf_oxy[0] = (ST)src.apf_oxy[0]; dst.apf_oxy[1] = (ST)src.apf_oxy[1];
  }
  for (int i = tid; i < kMaxNx; i += nthreads) { dst.dist_scale[i] = (CT)src.dist_scale[i]; dst.goal[i] = (ST)src.goal[i]; }
  for (int i = tid; i < kMaxNx * kMaxNx; i += nthreads) dst.Q[i] = (CT)src.Q[i];
  for (int i = tid; i < kMaxNu * kMaxNu; i += nthreads) dst.R[i] = (CT)src.R[i];
  for (int i = tid; i < TAMPC_MAX_TRAPS * kMaxNx; i += nthreads) dst.trap_x[i / kMaxNx][i % kMaxNx] = (ST)src.trap_x[i / kMaxNx][i % kMaxNx];
  for (int i = tid; i < TAMPC_MAX_TRAPS * kMaxNu; i += nthreads) dst.trap_u[i / kMaxNu][i % kMaxNu] = (CT)src.trap_u[i / kMaxNu][i % kMaxNu];
  for (int p = tid; p < TAMPC_MAX_TRAPS; p += nthreads) {
    CT n = (CT)0;
    for (int i = 0; i < kMaxNu; ++i)
      if ((src.usim_mask >> i) & 1u) { CT v = (CT)src.trap_u[p][i]; n = t_fma(v, v, n); }
    dst.trap_u_norm[p] = t_max(t_sqrt(n), (CT)kCosSimEps);
  }
  for (int i = tid; i < TAMPC_MAX_GOAL_SETS * TAMPC_MAX_SET_GOALS * kMaxNx; i += nthreads) {
    int j = i / (TAMPC_MAX_SET_GOALS * kMaxNx), g = (i / kMaxNx) % TAMPC_MAX_SET_GOALS, d = i % kMaxNx;
    dst.set_goals[j][g][d] = (ST)src.set_goals[j][g][d];
  }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory image.  Everything a rollout CTA needs that does not change between commands — packed network
// weights, CostS and SamplerS already converted to the kernel's types — sits contiguously in global memory in
// exactly the shared-memory layout (prep_image_kernel, run on set_model / set_cost), so a CTA's prologue is one
// flat asynchronous copy (LDGSTS, 16 bytes per request, all requests in flight at once) instead of dozens of
// dependent global round trips.
__device__ __forceinline__ void copy_image_async(unsigned char* smem_dst, const void* __restrict__ gsrc, int bytes, int tid, int nthreads) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned char* src = reinterpret_cast<const unsigned char*>(gsrc);
  for (int off = tid * 16; off < bytes; off += nthreads * 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + off), "l"(src + off) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void copy_image_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// sampler_d_off >= 0: additionally a float64 SamplerS (the tensor-core kernels' fused softmax update uses it)
template <typename ST, typename CT>
__global__ void prep_image_kernel(unsigned char* image, int cost_off, int sampler_off, int sampler_d_off, const tampc_cost_desc* cost,
                                  const SamplerD* sampler) {
  stage_cost<ST, CT>(*reinterpret_cast<CostS<ST, CT>*>(image + cost_off), *cost, threadIdx.x, blockDim.x);
  stage_sampler<ST>(*reinterpret_cast<SamplerS<ST>*>(image + sampler_off), *sampler, threadIdx.x, blockDim.x);
  if (sampler_d_off >= 0) stage_sampler<double>(*reinterpret_cast<SamplerS<double>*>(image + sampler_d_off), *sampler, threadIdx.x, blockDim.x);
}

}  // namespace tampc
