// fused_reduce.cuh — softmax update folded into the tail of the small-K rollout kernels (one launch per command).
//
// Same algebra and the same tampc_partial records as reduce_kernels.cuh (tail of pytorch_mppi `MPPI.command`):
// every warp reduces its own samples to a partial relative to its local minimum; the last warp to finish (atomic
// ticket) merges all partials with the log-sum-exp rescale and either finalises U / action (and stores the result
// straight into mapped host memory) or writes this rank's merged partial for the multi-GPU all-gather.
// Summation orders are functions of the launch shape only, so results are deterministic.
#pragma once
#include "rollout_kernel.cuh"

namespace tampc {

// Per-warp partials of the fused path live in a compact structure-of-arrays (headers contiguous so that the merging
// warp's loads coalesce; weighted sums with stride T*nu), not in the 4 KB ABI record.
struct PartialHdr {
  double beta, eta;
  long long argmin, count;
};

__device__ __forceinline__ void min_arg_combine(double& b, long long& bi, double o, long long oi) {
  if (o < b || (o == b && oi < bi)) { b = o; bi = oi; }
}

// One warp's partial.  Every lane passes the total cost of sample row `row` (< SW); rows beyond K_local are invalid.
template <int NU, int SW>
__device__ __forceinline__ void warp_partial(const RolloutArgs& a, const SamplerS<double>& sp, double total, int row, int kbase, int lane,
                                             int slot) {
  const unsigned full = 0xffffffffu;
  const bool valid = kbase + row < a.K_local;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double c = valid ? total : inf;
  double b = c;
  long long bi = row;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_xor_sync(full, b, off);
    const long long oi = __shfl_xor_sync(full, bi, off);
    min_arg_combine(b, bi, o, oi);
  }
  const double li = a.sampler->lambda_inv;
  const double wk = valid ? exp(-li * (c - b)) : 0.0;
  double eta = 0.0;
#pragma unroll
  for (int r = 0; r < SW; ++r) eta += __shfl_sync(full, wk, r);  // rows in order
  // weighted post-clamp noise: lane owns steps t = lane, lane+32, ...; rows with an underflowed weight are skipped
  constexpr int SLOTS = TAMPC_MAX_HORIZON / 32;
  double acc[SLOTS][NU];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s)
#pragma unroll
    for (int i = 0; i < NU; ++i) acc[s][i] = 0.0;
  const int T = a.src.T;
#pragma unroll 1
  for (int r = 0; r < SW; ++r) {
    const double wr = __shfl_sync(full, wk, r);
    if (wr == 0.0) continue;  // warp-uniform
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int t = lane + 32 * s;
      if (t < T) {
        double Ut[NU], u[NU], npost[NU];
        const int ts = t + 1;
#pragma unroll
        for (int d = 0; d < NU; ++d) Ut[d] = ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d];
        sample_action<double, NU>(a.src, sp, Ut, kbase + r, t, u, npost);
#pragma unroll
        for (int i = 0; i < NU; ++i) acc[s][i] = __fma_rn(wr, npost[i], acc[s][i]);
      }
    }
  }
  double* ws = a.pwsum + (size_t)slot * (T * NU);
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int t = lane + 32 * s;
    if (t < T) {
#pragma unroll
      for (int i = 0; i < NU; ++i) ws[t * NU + i] = acc[s][i];
    }
  }
  if (lane == 0) {
    PartialHdr h;
    h.beta = b;
    h.eta = eta;
    h.argmin = (long long)a.src.sample_offset + kbase + bi;
    const int cnt = a.K_local - kbase;
    h.count = cnt < 0 ? 0 : (cnt > SW ? SW : cnt);
    a.phdr[slot] = h;
  }
}

// Executed by the last warp of the grid: merge n partials (see combine_kernel for the algebra).
template <int NU>
__device__ __forceinline__ void last_warp_combine(const RolloutArgs& a, int n, int lane) {
  const unsigned full = 0xffffffffu;
  const PartialHdr* part = a.phdr;
  const double li = a.sampler->lambda_inv;
  double b = __longlong_as_double(0x7ff0000000000000ll);
  long long bi = 0x7fffffffffffffffll;
  long long cnt = 0;
  // The partials were written by other SMs: every load is an L2 round trip, so loads are issued in batches of 8
  // (independent) before anything consumes them.
  for (int p0 = lane; p0 < n; p0 += 32 * 8) {
    double be[8];
    long long ar[8], ct[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = p0 + 32 * u;
      const bool ok = p < n;
      be[u] = ok ? __ldcg(&part[p].beta) : __longlong_as_double(0x7ff0000000000000ll);
      ar[u] = ok ? __ldcg(&part[p].argmin) : 0x7fffffffffffffffll;
      ct[u] = ok ? __ldcg(&part[p].count) : 0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      min_arg_combine(b, bi, be[u], ar[u]);
      cnt += ct[u];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_xor_sync(full, b, off);
    const long long oi = __shfl_xor_sync(full, bi, off);
    min_arg_combine(b, bi, o, oi);
    cnt += __shfl_xor_sync(full, cnt, off);
  }
  const int T = a.src.T, nw = T * NU;
  constexpr int SLOTS = TAMPC_MAX_HORIZON * TAMPC_MAX_NU / 32;
  double acc[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) acc[s] = 0.0;
  double eta = 0.0;  // lane-strided partial sums, then a fixed xor tree
  for (int base0 = 0; base0 < n; base0 += 32 * 4) {
    double be[4], et[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = base0 + 32 * u + lane;
      const bool ok = p < n;
      be[u] = ok ? __ldcg(&part[p].beta) : __longlong_as_double(0x7ff0000000000000ll);
      et[u] = ok ? __ldcg(&part[p].eta) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int base = base0 + 32 * u;
      const double sc = (base + lane < n) ? exp(-li * (be[u] - b)) : 0.0;
      eta = __fma_rn(sc, et[u], eta);
      // partials in index order, eight at a time so that their loads are in flight together; a group whose scales
      // all underflowed (the usual case with a peaked softmax) is skipped, a zero scale inside a live group adds 0
      const unsigned live = __ballot_sync(full, sc != 0.0);
#pragma unroll 1
      for (int j0 = 0; j0 < 32; j0 += 8) {
        if (((live >> j0) & 0xffu) == 0u) continue;
        double sj[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) sj[q] = __shfl_sync(full, sc, j0 + q);
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
          const int i = lane + 32 * k;
          if (i < nw) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = (base + j0 + q < n) ? __ldcg(a.pwsum + (size_t)(base + j0 + q) * nw + i) : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[k] = __fma_rn(sj[q], v[q], acc[k]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) eta += __shfl_xor_sync(full, eta, off);
  if (a.fuse == 1) {
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int i = lane + 32 * k;
      if (i < nw) {
        const int t = i / NU, d = i % NU, ts = t + 1;
        const double base = ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d];
        const double v = base + acc[k] / eta;
        a.U_out[i] = v;
        if (t == 0) {
          a.action_out[d] = v;
          if (a.host_ll != nullptr) ll_publish(a.host_ll, d, v, a.seq);
        }
      }
    }
    if (lane == 0) {
      a.stats_out[0] = b;
      a.stats_out[1] = eta;
      a.stats_out[2] = (double)bi;
      if (a.host_ll != nullptr) {
        ll_publish(a.host_ll, kMaxNu, b, a.seq);
        ll_publish(a.host_ll, kMaxNu + 1, eta, a.seq);
        ll_publish(a.host_ll, kMaxNu + 2, (double)bi, a.seq);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int i = lane + 32 * k;
      if (i < nw) a.partial_out->wsum[i] = acc[k];
    }
    if (lane == 0) {
      a.partial_out->beta = b;
      a.partial_out->eta = eta;
      a.partial_out->argmin = bi;
      a.partial_out->count = cnt;
    }
  }
  if (lane == 0) *a.ticket = 0u;  // ready for the next launch on this stream
}

// tail of a rollout warp: partial, ticket, and the merge if this warp is the last one
template <int NU, int SW>
__device__ __forceinline__ void fused_tail(const RolloutArgs& a, const SamplerS<double>& sp, double total, int row, int kbase, int lane,
                                           int warp_global, int total_warps) {
  warp_partial<NU, SW>(a, sp, total, row, kbase, lane, warp_global);
  __threadfence();
  __syncwarp();
  unsigned ticket = 0;
  if (lane == 0) ticket = atomicAdd(a.ticket, 1u);
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket == (unsigned)(total_warps - 1)) {
    __threadfence();
    last_warp_combine<NU>(a, total_warps, lane);
  }
}

}  // namespace tampc
