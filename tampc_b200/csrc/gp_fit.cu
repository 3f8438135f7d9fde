// gp_fit.cu — hyper-parameter fit of the online GP residual on the device.
//
// Replaces OnlineGPMixing._fit_params (tampc/dynamics/online_model.py:412-430): `iters` Adam steps (lr 0.1, torch
// defaults) on the negative exact marginal log-likelihood of the independent-output GP
//     K_d = s_d exp(-|x_i - x_j|^2 / (2 l_d^2)) + (t_d + n) I,     loss = (1/n_out) sum_d 0.5 (r_d^T K_d^-1 r_d + log|K_d| + N log 2 pi)
// over the raw (softplus-constrained) parameters: lengthscale l_d, output scale s_d, task noise t_d (both noises under
// GreaterThan(1e-4)) and the global noise n shared by all outputs — gpytorch's parameterisation of
// ScaleKernel(RBFKernel) + MultitaskGaussianLikelihood (restated in oracle/shim/gpytorch; gpytorch itself is not under
// /root/reference).  The reference runs this on the host through autograd 15 times per control step (150 at a reset):
// here one CTA does it — warp d owns output d, everything float64 in shared memory — with the analytic gradient
//     d loss / d theta = (1/n_out) 0.5 ( tr(K^-1 dK/dtheta) - alpha^T dK/dtheta alpha ),   alpha = K^-1 r.
// The residual targets r = y - m(X) (mean function = nominal network at the training inputs) are constant during a fit.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/tampc_mppi.h"

namespace {

constexpr int kFitMaxPoints = TAMPC_GP_FIT_MAX_POINTS;  // 50: OnlineGPMixing.max_data_points (online_model.py:327)
constexpr int kFitMaxOut = TAMPC_GP_MAX_DIM;

__device__ __forceinline__ double softplus(double x) { return x > 20.0 ? x : log1p(exp(x)); }   // torch.nn.functional.softplus
__device__ __forceinline__ double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

struct FitArgs {
  int n, n_in, n_out, iters;
  const double* X;      // (n, n_in)
  const double* resid;  // (n, n_out)
  double* raw;          // [3 * n_out + 1]: raw lengthscale[n_out] | raw outputscale[n_out] | raw task noise[n_out] | raw noise
  double* adam_m;       // same layout
  double* adam_v;
  long long* step;      // Adam step count (shared by all parameters: they are always stepped together)
  double lr, beta1, beta2, eps;
  double* loss_out;     // [2]: loss of the last iteration, loss of the one before (last_loss_diff)
};

// one Adam step of a scalar (torch.optim.Adam, no weight decay / amsgrad)
__device__ __forceinline__ void adam_step(double& p, double& m, double& v, double g, long long t, const FitArgs& a) {
  m = a.beta1 * m + (1.0 - a.beta1) * g;
  v = a.beta2 * v + (1.0 - a.beta2) * g * g;
  const double bc1 = 1.0 - pow(a.beta1, (double)t), bc2 = 1.0 - pow(a.beta2, (double)t);
  const double denom = sqrt(v) / sqrt(bc2) + a.eps;
  p -= (a.lr / bc1) * (m / denom);
}

__global__ void __launch_bounds__(32 * kFitMaxOut) gp_fit_kernel(const FitArgs a) {
  extern __shared__ double sm[];
  const int n = a.n, ny = a.n_out;
  const int tid = threadIdx.x, lane = tid & 31, d = tid >> 5;
  double* D2 = sm;                                   // [n][n] squared distances, shared by all outputs
  double* V = D2 + n * n + (size_t)d * (n * n + 3 * n);  // this warp's matrix: K -> L -> L^-1 (in place)
  double* r = V + n * n;                             // residual targets of output d
  double* al = r + n;                                // w = L^-1 r, then alpha = K^-1 r
  double* xt = al + n;                               // scratch column
  __shared__ double s_gn[kFitMaxOut], s_nll[kFitMaxOut];

  for (int e = tid; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    double s = 0.0;
    for (int c = 0; c < a.n_in; ++c) { const double q = a.X[i * a.n_in + c] - a.X[j * a.n_in + c]; s += q * q; }
    D2[e] = s;
  }
  for (int i = lane; i < n; i += 32) r[i] = a.resid[i * ny + d];
  double rl = a.raw[d], ro = a.raw[ny + d], rt = a.raw[2 * ny + d], rn = a.raw[3 * ny];
  double m_l = a.adam_m[d], m_o = a.adam_m[ny + d], m_t = a.adam_m[2 * ny + d], m_n = a.adam_m[3 * ny];
  double v_l = a.adam_v[d], v_o = a.adam_v[ny + d], v_t = a.adam_v[2 * ny + d], v_n = a.adam_v[3 * ny];
  long long step = *a.step;
  double loss = 0.0, prev_loss = 0.0;
  __syncthreads();

  for (int it = 0; it < a.iters; ++it) {
    const double l = softplus(rl), s = softplus(ro), sig2 = softplus(rt) + 1e-4 + softplus(rn) + 1e-4;
    const double nh = -0.5 / (l * l);
    // ---- K (lower triangle is what the factorisation reads)
    for (int i = 0; i < n; ++i)  // (no runtime integer divisions in the loops of the fit: they cost more than the math)
      for (int j = lane; j <= i; j += 32) V[i * n + j] = s * exp(nh * D2[i * n + j]) + (i == j ? sig2 : 0.0);
    __syncwarp();
    // ---- Cholesky, in place, lower (right-looking)
    double logdet = 0.0;
    for (int j = 0; j < n; ++j) {
      const double piv = sqrt(V[j * n + j]);
      logdet += 2.0 * log(piv);
      __syncwarp();
      for (int i = j + lane; i < n; i += 32) V[i * n + j] = (i == j) ? piv : V[i * n + j] / piv;
      __syncwarp();
      const int rem = n - j - 1;  // trailing update of the lower triangle: V[i][k] -= V[i][j] V[k][j], j < k <= i
      for (int i0 = 0; i0 < rem; i0 += 4)   // 4 x 8 tiles of the trailing block, lower triangle only
        for (int k0 = 0; k0 <= i0 + 3 && k0 < rem; k0 += 8) {
          const int i = j + 1 + i0 + (lane >> 3), k = j + 1 + k0 + (lane & 7);
          if (i < n && k <= i) V[i * n + k] -= V[i * n + j] * V[k * n + j];
        }
      __syncwarp();
    }
    // ---- w = L^-1 r (forward), alpha = L^-T w (backward)
    for (int i = 0; i < n; ++i) {
      double p = 0.0;
      for (int k = lane; k < i; k += 32) p += V[i * n + k] * al[k];
      p = warp_sum(p);
      if (lane == 0) al[i] = (r[i] - p) / V[i * n + i];
      __syncwarp();
    }
    double quad = 0.0;
    for (int i = lane; i < n; i += 32) quad += al[i] * al[i];  // r^T K^-1 r = |L^-1 r|^2
    quad = warp_sum(quad);
    for (int i = n - 1; i >= 0; --i) {
      double p = 0.0;
      for (int k = i + 1 + lane; k < n; k += 32) p += V[k * n + i] * al[k];
      p = warp_sum(p);
      if (lane == 0) al[i] = (al[i] - p) / V[i * n + i];
      __syncwarp();
    }
    // ---- V = L^-1 in place (unblocked triangular inversion, columns from the last to the first)
    for (int j = n - 1; j >= 0; --j) {
      const double ajj = 1.0 / V[j * n + j];
      for (int i = j + 1 + lane; i < n; i += 32) xt[i] = V[i * n + j];
      __syncwarp();
      for (int i = j + 1 + lane; i < n; i += 32) {  // x <- T x with T = the already inverted trailing block
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;  // four partial sums: the dot product is latency-bound otherwise
        int k = j + 1;
        for (; k + 3 <= i; k += 4) {
          p0 += V[i * n + k] * xt[k];
          p1 += V[i * n + k + 1] * xt[k + 1];
          p2 += V[i * n + k + 2] * xt[k + 2];
          p3 += V[i * n + k + 3] * xt[k + 3];
        }
        for (; k <= i; ++k) p0 += V[i * n + k] * xt[k];
        V[i * n + j] = -ajj * ((p0 + p1) + (p2 + p3));
      }
      if (lane == 0) V[j * n + j] = ajj;
      __syncwarp();
    }
    // ---- traces and quadratic forms over the lower triangle (K^-1 = V^T V; E = exp(..), G = E * D2)
    double trE = 0.0, trG = 0.0, trI = 0.0, aEa = 0.0, aGa = 0.0, aa = 0.0;
    for (int i = 0; i < n; ++i)
    for (int j = lane; j <= i; j += 32) {
      const int e = i * n + j;
      double k0 = 0.0, k1 = 0.0, k2 = 0.0, k3 = 0.0;
      int k = i;
      for (; k + 3 < n; k += 4) {
        k0 += V[k * n + i] * V[k * n + j];
        k1 += V[(k + 1) * n + i] * V[(k + 1) * n + j];
        k2 += V[(k + 2) * n + i] * V[(k + 2) * n + j];
        k3 += V[(k + 3) * n + i] * V[(k + 3) * n + j];
      }
      for (; k < n; ++k) k0 += V[k * n + i] * V[k * n + j];
      const double kinv = (k0 + k1) + (k2 + k3);
      const double w = (i == j) ? 1.0 : 2.0;
      const double E = exp(nh * D2[e]), G = E * D2[e];
      trE += w * kinv * E;
      trG += w * kinv * G;
      aEa += w * al[i] * al[j] * E;
      aGa += w * al[i] * al[j] * G;
      if (i == j) { trI += kinv; aa += al[i] * al[i]; }
    }
    trE = warp_sum(trE); trG = warp_sum(trG); trI = warp_sum(trI);
    aEa = warp_sum(aEa); aGa = warp_sum(aGa); aa = warp_sum(aa);
    const double nll = 0.5 * (quad + logdet + n * 1.8378770664093453);  // log(2 pi)
    const double inv_ny = 1.0 / ny;
    const double g_s = 0.5 * (trE - aEa), g_l = 0.5 * (s / (l * l * l)) * (trG - aGa), g_sig = 0.5 * (trI - aa);
    if (lane == 0) { s_gn[d] = g_sig; s_nll[d] = nll; }
    __syncthreads();
    double gsum = 0.0, lsum = 0.0;
    for (int q = 0; q < ny; ++q) { gsum += s_gn[q]; lsum += s_nll[q]; }
    prev_loss = loss;
    loss = lsum * inv_ny;
    ++step;
    adam_step(ro, m_o, v_o, inv_ny * g_s * sigmoid(ro), step, a);
    adam_step(rl, m_l, v_l, inv_ny * g_l * sigmoid(rl), step, a);
    adam_step(rt, m_t, v_t, inv_ny * g_sig * sigmoid(rt), step, a);
    adam_step(rn, m_n, v_n, inv_ny * gsum * sigmoid(rn), step, a);  // every thread keeps an identical copy
    __syncthreads();
  }
  if (lane == 0) {
    a.raw[d] = rl; a.raw[ny + d] = ro; a.raw[2 * ny + d] = rt;
    a.adam_m[d] = m_l; a.adam_m[ny + d] = m_o; a.adam_m[2 * ny + d] = m_t;
    a.adam_v[d] = v_l; a.adam_v[ny + d] = v_o; a.adam_v[2 * ny + d] = v_t;
    if (d == 0) {
      a.raw[3 * ny] = rn; a.adam_m[3 * ny] = m_n; a.adam_v[3 * ny] = v_n;
      *a.step = step;
      a.loss_out[0] = loss;
      a.loss_out[1] = prev_loss;
    }
  }
}

thread_local char g_fit_error[256] = "";

}  // namespace

extern "C" {

TAMPC_API const char* tampc_gp_fit_last_error(void) { return g_fit_error; }

TAMPC_API int tampc_gp_fit(int32_t device, int32_t n_points, int32_t n_in, int32_t n_out, const double* X, const double* resid, int32_t iters,
                           double lr, double* raw_params, double* adam_m, double* adam_v, int64_t* adam_step, double* loss_out) {
  auto fail = [&](int code, const char* msg, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(g_fit_error, sizeof g_fit_error, "%s: %s", msg, cudaGetErrorString(e));
    else snprintf(g_fit_error, sizeof g_fit_error, "%s", msg);
    return code;
  };
  if (!X || !resid || !raw_params || !adam_m || !adam_v || !adam_step || !loss_out) return fail(TAMPC_ERR_INVALID, "null argument");
  if (n_points < 1 || n_points > kFitMaxPoints) return fail(TAMPC_ERR_UNSUPPORTED, "GP fit supports 1..TAMPC_GP_FIT_MAX_POINTS data points");
  if (n_in < 1 || n_in > TAMPC_GP_MAX_DIM || n_out < 1 || n_out > kFitMaxOut) return fail(TAMPC_ERR_UNSUPPORTED, "GP fit: input/output width out of range");
  if (iters < 0) return fail(TAMPC_ERR_INVALID, "negative iteration count");
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(TAMPC_ERR_NO_DEVICE, "cudaSetDevice", e);
  const int np = 3 * n_out + 1;
  const size_t bytes_in = sizeof(double) * ((size_t)n_points * n_in + (size_t)n_points * n_out);
  const size_t bytes_st = sizeof(double) * (3 * (size_t)np + 2) + sizeof(long long);
  // one small scratch buffer per device, kept for the life of the process (a fit runs every control step)
  constexpr size_t kScratch = sizeof(double) * (2 * (size_t)kFitMaxPoints * TAMPC_GP_MAX_DIM + 3 * (3 * kFitMaxOut + 1) + 2) + 64;
  static unsigned char* g_scratch[64] = {nullptr};
  if (device < 0 || device >= 64) return fail(TAMPC_ERR_INVALID, "device ordinal out of range");
  if (!g_scratch[device] && (e = cudaMalloc(&g_scratch[device], kScratch)) != cudaSuccess) return fail(TAMPC_ERR_CUDA, "cudaMalloc", e);
  unsigned char* dbuf = g_scratch[device];
  static_assert(kScratch >= 8, "scratch");
  if (bytes_in + bytes_st + 64 > kScratch) return fail(TAMPC_ERR_INVALID, "internal: scratch too small");
  double* dX = reinterpret_cast<double*>(dbuf);
  double* dR = dX + (size_t)n_points * n_in;
  double* draw = dR + (size_t)n_points * n_out;
  double* dm = draw + np;
  double* dv = dm + np;
  double* dloss = dv + np;
  long long* dstep = reinterpret_cast<long long*>(dloss + 2);
  long long step = *adam_step;
  bool ok = true;
  ok = ok && (e = cudaMemcpy(dX, X, sizeof(double) * n_points * n_in, cudaMemcpyHostToDevice)) == cudaSuccess;
  ok = ok && (e = cudaMemcpy(dR, resid, sizeof(double) * n_points * n_out, cudaMemcpyHostToDevice)) == cudaSuccess;
  ok = ok && (e = cudaMemcpy(draw, raw_params, sizeof(double) * np, cudaMemcpyHostToDevice)) == cudaSuccess;
  ok = ok && (e = cudaMemcpy(dm, adam_m, sizeof(double) * np, cudaMemcpyHostToDevice)) == cudaSuccess;
  ok = ok && (e = cudaMemcpy(dv, adam_v, sizeof(double) * np, cudaMemcpyHostToDevice)) == cudaSuccess;
  ok = ok && (e = cudaMemcpy(dstep, &step, sizeof step, cudaMemcpyHostToDevice)) == cudaSuccess;
  if (ok) {
    FitArgs a{};
    a.n = n_points; a.n_in = n_in; a.n_out = n_out; a.iters = iters;
    a.X = dX; a.resid = dR; a.raw = draw; a.adam_m = dm; a.adam_v = dv; a.step = dstep;
    a.lr = lr; a.beta1 = 0.9; a.beta2 = 0.999; a.eps = 1e-8;
    a.loss_out = dloss;
    const size_t smem = sizeof(double) * ((size_t)n_points * n_points + (size_t)n_out * ((size_t)n_points * n_points + 3 * n_points));
    ok = (e = cudaFuncSetAttribute(gp_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) == cudaSuccess;
    if (ok) {
      gp_fit_kernel<<<1, 32 * n_out, smem>>>(a);
      ok = (e = cudaGetLastError()) == cudaSuccess && (e = cudaDeviceSynchronize()) == cudaSuccess;
    }
  }
  ok = ok && (e = cudaMemcpy(raw_params, draw, sizeof(double) * np, cudaMemcpyDeviceToHost)) == cudaSuccess;
  ok = ok && (e = cudaMemcpy(adam_m, dm, sizeof(double) * np, cudaMemcpyDeviceToHost)) == cudaSuccess;
  ok = ok && (e = cudaMemcpy(adam_v, dv, sizeof(double) * np, cudaMemcpyDeviceToHost)) == cudaSuccess;
  ok = ok && (e = cudaMemcpy(&step, dstep, sizeof step, cudaMemcpyDeviceToHost)) == cudaSuccess;
  ok = ok && (e = cudaMemcpy(loss_out, dloss, sizeof(double) * 2, cudaMemcpyDeviceToHost)) == cudaSuccess;
  if (!ok) return fail(TAMPC_ERR_CUDA, "GP fit", e);
  *adam_step = step;
  return TAMPC_OK;
}

}  // extern "C"
