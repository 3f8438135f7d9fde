// gp_rollout.cuh — fused rollout with the online GP residual.
//
// Replaces, per sample-step, OnlineDynamicsModel.predict -> OnlineGPMixing._make_prediction / sample
// (tampc/dynamics/online_model.py:67-99, 435-449; MixingBatchGP :286-315): for every output d an exact-GP posterior
//     mean_d = m_d(g) + k_d(g, X) alpha_d                   alpha_d = (K_d + n_d I)^-1 (y_d - m_d(X))
//     var_d  = s_d + n_d - | L_d^-1 k_d(g, X) |^2           K_d + n_d I = L_d L_d^T,  k_d(a,b) = s_d exp(-|a-b|^2 / (2 l_d^2))
// followed by a Normal(mean, sqrt(var)) draw.  The mean function m is the nominal network, which the kernel evaluates
// anyway; alpha and L^-1 are prepared on the host once per GP update (tampc_mppi_set_gp), staged in shared memory and
// read with broadcast loads.  The posterior is written with L^-1 (a sum of squares) rather than (K + n I)^-1.
// The GP part — kernel vector, mean, variance, draw — is evaluated in float64 (type GT) in EVERY precision mode: alpha
// has large entries of alternating sign (the kernel matrix of <= 50 nearby points is ill-conditioned), so float32
// rounding of the individual k_j is amplified by |alpha| in the mean, and the variance is a difference of O(s) terms;
// with float32 here the per-sample costs only reached 5e-5 relative (north star: 1e-5).  The networks around it stay in
// the mode's arithmetic (float32 FFMA2 / float64 DFMA).
//
// A stochastic model makes the reference's M rollout replicas (ExperimentalMPPI._compute_rollout_costs,
// tampc/controller/controller.py:340-381) genuinely different: one thread per (sample, replica), the replicas of a
// sample in adjacent lanes of one warp, mean / unbiased variance across replicas by warp shuffles.
#pragma once
#include "rollout_kernel.cuh"

namespace tampc {

constexpr int kGpMaxPoints = TAMPC_GP_MAX_POINTS;
constexpr int kGpMaxDim = TAMPC_GP_MAX_DIM;

// Shared-memory block: header | X[npad][8] | alpha[n_out][npad] | tri[n_out][tri_stride]   (all of type T)
// tri = L^-1 row by row, row i holding entries 0..i padded with zeros to a multiple of 4.
template <typename T>
struct GpHeader {
  int n, npad, n_in, n_out, sample, space, tri_stride, pad_;
  T in_scale[kGpMaxDim], in_off[kGpMaxDim];  // g = input * in_scale + in_off
  T out_mul[kGpMaxDim];                       // what is added to the model's output: delta_d * out_mul_d
  T neg_half_inv_l2[kGpMaxDim];               // -1 / (2 l_d^2)
  T oscale[kGpMaxDim];                        // s_d
  T prior_var[kGpMaxDim];                     // s_d + n_d
};

__host__ __device__ constexpr int gp_tri_stride(int n) {
  int s = 0;
  for (int i = 0; i < n; ++i) s += (i + 4) & ~3;
  return s;
}
template <typename T> __host__ __device__ constexpr size_t gp_block_bytes(int n, int n_out) {
  const int npad = (n + 3) & ~3;
  return (sizeof(GpHeader<T>) + sizeof(T) * ((size_t)npad * kGpMaxDim + (size_t)n_out * npad + (size_t)n_out * gp_tri_stride(n)) + 15) & ~size_t(15);
}

__device__ __forceinline__ float t_exp(float a) { return expf(a); }
__device__ __forceinline__ double t_exp(double a) { return exp(a); }

// four consecutive elements from a 16-byte aligned address (one LDS.128 for float, two for double)
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
__device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
  const double2 q0 = *reinterpret_cast<const double2*>(p), q1 = *reinterpret_cast<const double2*>(p + 2);
  v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
}

// Four-element vectors of the GP's arithmetic type (shared-memory staging of the kernel vector k(g, X)).
template <typename T> struct GpVec4;
template <> struct GpVec4<float> { using type = float4; };
template <> struct GpVec4<double> { using type = double4; };
template <typename T> __host__ __device__ constexpr size_t gp_scratch_bytes(int threads) { return sizeof(T) * kGpMaxPoints * (size_t)threads; }

// delta[d] = (k_d alpha_d + sqrt(var_d) eps_d) * out_mul_d  for the input row `in` (NIN entries).
// `ks`: this thread's slot of the per-CTA scratch, 4-element vectors `kstride` apart (layout [j/4][thread]): the
// kernel vector of one output lives there between its evaluation and the triangular product.  (A per-thread local
// array would live in L1, most of which the shared-memory carve-out takes: measured 144 cycles per 4 FMAs.)
template <typename T, int NIN, int NOUT>
__device__ __forceinline__ void gp_delta(const unsigned char* __restrict__ gp, typename GpVec4<T>::type* __restrict__ ks, int kstride,
                                         const T (&in)[NIN], const T (&eps)[NOUT], T (&delta)[NOUT]) {
  using V4 = typename GpVec4<T>::type;
  const GpHeader<T>& H = *reinterpret_cast<const GpHeader<T>*>(gp);
  const T* __restrict__ X = reinterpret_cast<const T*>(gp + sizeof(GpHeader<T>));
  const int n = H.n, npad = H.npad;
  const T* __restrict__ alpha = X + npad * kGpMaxDim;
  const T* __restrict__ tri = alpha + NOUT * npad;
  T g[NIN];
#pragma unroll
  for (int i = 0; i < NIN; ++i) g[i] = t_fma(in[i], H.in_scale[i], H.in_off[i]);
#pragma unroll 1
  for (int d = 0; d < NOUT; ++d) {
    T mean = (T)0;
    const T nh = H.neg_half_inv_l2[d], os = H.oscale[d];
#pragma unroll 2
    for (int j4 = 0; j4 < npad; j4 += 4) {
      T kq[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j4 + e;
        T s = (T)0;
#pragma unroll
        for (int i = 0; i < NIN; ++i) {
          const T dd = g[i] - X[j * kGpMaxDim + i];
          s = t_fma(dd, dd, s);
        }
        kq[e] = j < n ? os * t_exp(nh * s) : (T)0;
        mean = t_fma(alpha[d * npad + j], kq[e], mean);
      }
      V4 v;
      v.x = kq[0]; v.y = kq[1]; v.z = kq[2]; v.w = kq[3];
      ks[(j4 >> 2) * kstride] = v;
    }
    // q = | L^-1 k |^2, row by row over the packed lower triangle: one broadcast load of four matrix entries and
    // one load of four kernel values per four FMAs, unrolled so that several loads are in flight
    T q = (T)0;
    const T* __restrict__ row = tri + (size_t)d * H.tri_stride;
    for (int i = 0; i < n; ++i) {
      const int len = (i + 4) & ~3;
      T acc[4] = {(T)0, (T)0, (T)0, (T)0};
#pragma unroll 4
      for (int j = 0; j < len; j += 4) {
        T r[4];
        load4(row + j, r);
        const V4 kk = ks[(j >> 2) * kstride];
        acc[0] = t_fma(r[0], kk.x, acc[0]);
        acc[1] = t_fma(r[1], kk.y, acc[1]);
        acc[2] = t_fma(r[2], kk.z, acc[2]);
        acc[3] = t_fma(r[3], kk.w, acc[3]);
      }
      const T w = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      q = t_fma(w, w, q);
      row += len;
    }
    // exact arithmetic keeps var >= n_d > 0; the clamp only guards float32 rounding
    const T var = t_max(H.prior_var[d] - q, (T)0);
    T out = mean;
    if (H.sample) out = t_fma(t_sqrt(var), eps[d], out);
    delta[d] = out * H.out_mul[d];
  }
}

// ---- models with the residual hook ------------------------------------------------------------------------
// space 0: GP input = the network's input row, output added to the network's (un-scaled) output;
// space 1: GP input = [x,u], output added to dx.  (For a plain MLP both are the same form, see mppi_api.cu.)
template <class Model> struct GpModel;

template <typename NT, int NX_, int NU_, int H1, int H2>
struct GpModel<ModelMlp<NT, NX_, NU_, H1, H2>> {
  using Base = ModelMlp<NT, NX_, NU_, H1, H2>;
  static constexpr int NOUT = NX_;
  template <typename ST, typename GT>
  static __device__ __forceinline__ void predict(const NT* __restrict__ w, NT slope, const unsigned char* __restrict__ gp,
                                                 typename GpVec4<GT>::type* __restrict__ ks, int kstride, const GT (&eps)[NOUT],
                                                 ST (&x)[NX_], const ST (&u)[NU_]) {
    NT in[1][NX_ + NU_];
    GT gin[NX_ + NU_];
#pragma unroll
    for (int i = 0; i < NX_; ++i) { in[0][i] = (NT)x[i]; gin[i] = (GT)x[i]; }
#pragma unroll
    for (int i = 0; i < NU_; ++i) { in[0][NX_ + i] = (NT)u[i]; gin[NX_ + i] = (GT)u[i]; }
    NT dx[1][NX_];
    Base::Net::template eval<1>(w, in, dx, slope);
    GT delta[NX_];
    gp_delta<GT, NX_ + NU_, NX_>(gp, ks, kstride, gin, eps, delta);
#pragma unroll
    for (int i = 0; i < NX_; ++i) x[i] += (ST)((GT)dx[0][i] + delta[i]);
  }
};

template <typename NT, int NX_, int NU_, int E1, int E2, int NZ, int D1, int D2, int NV, int C1, int C2>
struct GpModel<ModelRex<NT, NX_, NU_, E1, E2, NZ, D1, D2, NV, C1, C2>> {
  using Base = ModelRex<NT, NX_, NU_, E1, E2, NZ, D1, D2, NV, C1, C2>;
  static constexpr int NOUT = NX_ > NV ? NX_ : NV;
  template <typename ST, typename GT>
  static __device__ __forceinline__ void predict(const NT* __restrict__ w, NT slope, const unsigned char* __restrict__ gp,
                                                 typename GpVec4<GT>::type* __restrict__ ks, int kstride, const GT (&eps)[NOUT],
                                                 ST (&x)[NX_], const ST (&u)[NU_]) {
    const int space = reinterpret_cast<const GpHeader<GT>*>(gp)->space;
    NT in[1][NX_ + NU_];
#pragma unroll
    for (int i = 0; i < NX_; ++i) in[0][i] = (NT)x[i];
#pragma unroll
    for (int i = 0; i < NU_; ++i) in[0][NX_ + i] = (NT)u[i];
    GT v[NV];
    {
      NT z[1][NZ], vn[1][NV];
      Base::Enc::template eval<1>(w, in, z, slope);
      Base::Dyn::template eval<1>(w + Base::kOffDyn, z, vn, slope);
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = (GT)vn[0][j];
      if (space == TAMPC_GP_AT_NETWORK) {
        GT e[NV], delta[NV], zin[NZ];
#pragma unroll
        for (int j = 0; j < NV; ++j) e[j] = eps[j];
#pragma unroll
        for (int j = 0; j < NZ; ++j) zin[j] = (GT)z[0][j];
        gp_delta<GT, NZ, NV>(gp, ks, kstride, zin, e, delta);
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] += delta[j];
      }
    }
    NT xin[1][NX_];
#pragma unroll
    for (int i = 0; i < NX_; ++i) xin[0][i] = (NT)x[i];
    NT C[1][NX_ * NV];
    Base::Dec::template eval<1>(w + Base::kOffDec, xin, C, slope);
    GT dx[NX_];
#pragma unroll
    for (int i = 0; i < NX_; ++i) {
      GT acc = (GT)w[Base::kOffD0 + i];
#pragma unroll
      for (int j = 0; j < NV; ++j) acc = t_fma((GT)C[0][i * NV + j], v[j], acc);
      dx[i] = acc;
    }
    if (space == TAMPC_GP_ORIGINAL) {
      GT e[NX_], delta[NX_], gin[NX_ + NU_];
#pragma unroll
      for (int i = 0; i < NX_; ++i) { e[i] = eps[i]; gin[i] = (GT)x[i]; }
#pragma unroll
      for (int i = 0; i < NU_; ++i) gin[NX_ + i] = (GT)u[i];
      gp_delta<GT, NX_ + NU_, NX_>(gp, ks, kstride, gin, e, delta);
#pragma unroll
      for (int i = 0; i < NX_; ++i) dx[i] += delta[i];
    }
#pragma unroll
    for (int i = 0; i < NX_; ++i) x[i] += (ST)dx[i];
  }
};

// ---- kernel ---------------------------------------------------------------------------------------------
// One thread per (sample, replica).  A warp holds G = 32 / M samples; lanes G*M..31 mirror the last group and never
// write.  Shared memory: the FMA-path image | shifted nominal sequence | GP block | per-thread kernel-vector scratch.
template <class Model, typename NT, typename ST, typename CT, typename GT>
__global__ void __launch_bounds__(128) rollout_gp_kernel(const RolloutArgs a) {
  using GM = GpModel<Model>;
  constexpr int NX = Model::NX, NU = Model::NU, NOUT = GM::NOUT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = RolloutSmem<Model, NT, ST, CT>;
  NT* w = reinterpret_cast<NT*>(smem_raw + L::kWeights);
  CostS<ST, CT>& cs = *reinterpret_cast<CostS<ST, CT>*>(smem_raw + L::kCost);
  SamplerS<ST>& sp = *reinterpret_cast<SamplerS<ST>*>(smem_raw + L::kSampler);
  ST* Ush = reinterpret_cast<ST*>(smem_raw + L::kU);
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
  const int T = a.src.T;
  unsigned char* gp = smem_raw + L::bytes(T);
  typename GpVec4<GT>::type* ks = reinterpret_cast<typename GpVec4<GT>::type*>(gp + a.gp_bytes) + tid;  // [j/4][thread] scratch
  {
    copy_image_async(smem_raw, a.image, (int)L::kU, tid, nthr);
    copy_image_async(gp, a.gp, a.gp_bytes, tid, nthr);
    for (int i = tid; i < T * NU; i += nthr) {
      int t = i / NU, d = i % NU, ts = t + a.shift;
      Ush[t * kMaxNu + d] = (ST)(ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d]);
    }
    copy_image_wait();
  }
  __syncthreads();

  const int M = a.M, G = 32 / M;
  const int grp = min(lane / M, G - 1), base = grp * M, m = min(lane - base, M - 1);
  const bool active = lane < G * M;
  const int kk = (blockIdx.x * (nthr >> 5) + (tid >> 5)) * G + grp;
  const bool live = active && kk < a.K_local;
  const int k = min(kk, a.K_local - 1);
  const int ny = reinterpret_cast<const GpHeader<GT>*>(gp)->n_out;

  ST x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = (ST)a.state[i];
  ST cost = (ST)0, pert = (ST)0;
  double cvar = 0.0, disc = 1.0;
  const NT slope = (NT)a.slope;
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
    GT eps[NOUT];
    if (a.gp_eps != nullptr) {  // injected standard normals (K_local, M, T, n_out)
      const double* p = a.gp_eps + (((size_t)k * M + m) * T + t) * ny;
#pragma unroll
      for (int i = 0; i < NOUT; ++i) eps[i] = i < ny ? (GT)p[i] : (GT)0;
    } else {  // Philox stream of (global sample, replica, t): counter word 1 = 128 * (1 + 2 m + block) + t
      float z[8];
      const uint32_t s = (uint32_t)(a.src.sample_offset + k);
      philox_normal4(a.src.seed, a.src.call, s, 128u * (1u + 2u * (uint32_t)m) + (uint32_t)t, *reinterpret_cast<float(*)[4]>(z));
      if (NOUT > 4) philox_normal4(a.src.seed, a.src.call, s, 128u * (2u + 2u * (uint32_t)m) + (uint32_t)t, *reinterpret_cast<float(*)[4]>(z + 4));
#pragma unroll
      for (int i = 0; i < NOUT; ++i) eps[i] = (GT)z[i];
    }
    GM::template predict<ST, GT>(w, slope, gp, ks, nthr, eps, x, u);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      if ((a.wrap_mask >> i) & 1u) x[i] = angle_normalize<ST>(x[i]);
      if (bad) x[i] = (ST)0;
    }
    const CT c = running_cost<ST, CT, NX, NU>(cs, x, u);
    cost += (ST)c;
    if (a.var_cost != 0.0 && M > 1) {
      // c.var(dim=0) over the replicas (controller.py:364): unbiased, about the mean
      const double cd = (double)c;
      double s = 0.0;
      for (int j = 0; j < M; ++j) s += __shfl_sync(0xffffffffu, cd, base + j);
      const double mean = s / M;
      double ss = 0.0;
      for (int j = 0; j < M; ++j) {
        const double dv = __shfl_sync(0xffffffffu, cd, base + j) - mean;
        ss = __fma_rn(dv, dv, ss);
      }
      cvar += ss / (M - 1) * disc;
      disc *= a.var_discount;
    }
    if (a.states_out != nullptr && live && m == 0) {
      double* so = a.states_out + ((size_t)k * T + t) * NX;
#pragma unroll
      for (int i = 0; i < NX; ++i) so[i] = (double)x[i];
    }
  }
  const double mine = (double)(cost + (ST)terminal_cost<ST, CT, NX>(cs, x));
  double s = 0.0;
  for (int j = 0; j < M; ++j) s += __shfl_sync(0xffffffffu, mine, base + j);  // cost_samples.mean(dim=0), fixed order
  const double total = s / M + cvar * a.var_cost + (double)pert;
  if (live && m == 0) a.cost_out[k] = total;
}

template <class Model, typename NT, typename ST, typename CT, typename GT>
cudaError_t launch_rollout_gp(const RolloutArgs& a, int block, cudaStream_t stream) {
  auto kern = rollout_gp_kernel<Model, NT, ST, CT, GT>;
  const size_t fixed = RolloutSmem<Model, NT, ST, CT>::bytes(a.src.T) + (size_t)a.gp_bytes;
  while (block > 32 && fixed + gp_scratch_bytes<GT>(block) > 227 * 1024) block >>= 1;  // many points: smaller CTAs
  const size_t smem = fixed + gp_scratch_bytes<GT>(block);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int per_block = (block / 32) * (32 / a.M);
  const int grid = (a.K_local + per_block - 1) / per_block;
  kern<<<grid, block, smem, stream>>>(a);
  return cudaGetLastError();
}

struct GpLaunchers {
  rollout_launch_fn f64 = nullptr, mixed = nullptr, f32 = nullptr;
  rollout_launch_fn nominal = nullptr;  // plain FMA kernel over the SAME image (float32 networks, float64 state + cost): predict_nominal
};
// Every non-float64 mode runs float32 networks with the state, the cost terms and the GP part in float64: with
// rollout_var_cost > 0 the cost holds the VARIANCE of the replicas' step costs (controller.py:364), a difference of nearly
// equal numbers that float32 cost terms resolve to ~1e-3 only (measured 6e-5 on the total, north star 1e-5).
template <class MF, class MD>
GpLaunchers make_gp_launchers() {
  GpLaunchers g;
  g.f64 = &launch_rollout_gp<MD, double, double, double, double>;
  g.mixed = &launch_rollout_gp<MF, float, double, double, double>;
  g.f32 = g.mixed;
  g.nominal = &launch_rollout<MF, float, double, double, 1>;
  return g;
}

}  // namespace tampc
