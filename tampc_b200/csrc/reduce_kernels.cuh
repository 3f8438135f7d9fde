// reduce_kernels.cuh — softmax weighting of the sampled trajectories and update of the nominal sequence.
//
// Replaces the tail of pytorch_mppi `MPPI.command` (restated in oracle/shim/pytorch_mppi/mppi.py; called from
// tampc/controller/online_controller.py:497):  beta = min cost;  w = exp(-(cost-beta)/lambda);  eta = sum w;
// U[t] += sum_k (w_k/eta) * noise[k,t]  with the POST-clamp noise of controller.py:394.
//
// Two stages so that one design serves one GPU and a K-sharded multi-GPU run:
//   partials_kernel : each CTA reduces a chunk of samples to a tampc_partial relative to its own local minimum
//   combine_kernel  : one CTA merges P partials (CTA partials, or one per rank after the all-gather) in index
//                     order with the log-sum-exp rescale s_p = exp(-(beta_p - beta)/lambda); deterministic, so
//                     every rank obtains bitwise-identical U.
// The post-clamp noise is regenerated (Philox) or re-read (injected) rather than stored by the rollout kernel.
#pragma once
#include "device_common.cuh"

namespace tampc {

struct ReduceArgs {
  const double* cost;   // (K_local,)
  const SamplerD* sampler;
  const double* U;      // nominal sequence BEFORE the shift (T, nu)
  SampleSrc src;
  int K_local, chunk, nu;
  tampc_partial* partials;  // (n_chunks,)
};

constexpr int kReduceThreads = 256;

// Block-wide (min, lowest index) and sum with warp shuffles and one shared-memory hop: two barriers instead of one per
// tree level.  The combination order is a function of the block shape only, so results are deterministic.
__device__ __forceinline__ void warp_min_arg(double& b, long long& bi) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, b, off);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (o < b || (o == b && oi < bi)) { b = o; bi = oi; }
  }
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
// s_b / s_i: one slot per warp (<= 32 warps).  Every thread returns the block's (min, argmin).
__device__ __forceinline__ void block_min_arg(double& b, long long& bi, double* s_b, long long* s_i, int tid, int nthreads) {
  warp_min_arg(b, bi);
  if ((tid & 31) == 0) { s_b[tid >> 5] = b; s_i[tid >> 5] = bi; }
  __syncthreads();
  const int nw = nthreads >> 5;
  b = (tid & 31) < nw ? s_b[tid & 31] : __longlong_as_double(0x7ff0000000000000ll);
  bi = (tid & 31) < nw ? s_i[tid & 31] : 0x7fffffffffffffffll;
  warp_min_arg(b, bi);
  __syncthreads();  // the slots may be reused
}
__device__ __forceinline__ double block_sum(double v, double* s_b, int tid, int nthreads) {
  v = warp_sum(v);
  if ((tid & 31) == 0) s_b[tid >> 5] = v;
  __syncthreads();
  const int nw = nthreads >> 5;
  v = warp_sum((tid & 31) < nw ? s_b[tid & 31] : 0.0);
  __syncthreads();
  return v;
}

// One CTA's partial; NT = block size (the chunk holds at most NT samples).
template <int NU, int NT>
__device__ __forceinline__ void partials_body(const ReduceArgs& a) {
  __shared__ double s_w[NT];      // weights of the chunk (chunk <= NT)
  __shared__ SamplerS<double> sp;
  __shared__ double Ush[TAMPC_MAX_HORIZON * kMaxNu];
  __shared__ double s_acc[NT * NU];

  const int tid = threadIdx.x;
  const int T = a.src.T;
  const int k0 = blockIdx.x * a.chunk;
  const int n = min(a.chunk, a.K_local - k0);
  stage_sampler<double>(sp, *a.sampler, tid, NT);
  for (int i = tid; i < T * NU; i += NT) {
    int t = i / NU, d = i % NU, ts = t + 1;
    Ush[t * kMaxNu + d] = ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d];
  }
  // ---- block min / argmin (lowest index wins ties, like torch.min / argmin over a 1-D tensor)
  __shared__ double s_wb[32];
  __shared__ long long s_wi[32];
  grid_dep_wait();  // the costs come from the rollout launch in front (everything staged above predates it)
  const double c = tid < n ? a.cost[k0 + tid] : __longlong_as_double(0x7ff0000000000000ll);
  double beta = c;
  long long amin_ll = tid;
  block_min_arg(beta, amin_ll, s_wb, s_wi, tid, NT);
  const int amin = (int)amin_ll;
  // ---- weights and their sum
  const double w = tid < n ? exp(-a.sampler->lambda_inv * (c - beta)) : 0.0;
  s_w[tid] = w;
  const double eta = block_sum(w, s_wb, tid, NT);  // (its barriers also publish s_w, sp and Ush)
  // ---- weighted post-clamp noise: thread (t, g) sums samples g, g+G, ... of the chunk for step t  (T <= 128)
  const int G = NT / T;
  if (tid < G * T) {
    const int t = tid % T, g = tid / T;
    double acc[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) acc[i] = 0.0;
    for (int j = g; j < n; j += G) {
      const double wj = s_w[j];
      if (wj == 0.0) continue;  // exp underflow: contributes exactly nothing
      double u[NU], npost[NU];
      sample_action<double, NU>(a.src, sp, Ush + t * kMaxNu, k0 + j, t, u, npost);
#pragma unroll
      for (int i = 0; i < NU; ++i) acc[i] = __fma_rn(wj, npost[i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) s_acc[tid * NU + i] = acc[i];
  }
  __syncthreads();
  tampc_partial& out = a.partials[blockIdx.x];
  for (int i = tid; i < T * NU; i += NT) {
    const int t = i / NU, d = i % NU;
    double sum = 0.0;
    for (int g = 0; g < G; ++g) sum += s_acc[(g * T + t) * NU + d];
    out.wsum[t * NU + d] = sum;
  }
  if (tid == 0) {
    out.beta = beta;
    out.eta = eta;
    out.argmin = (int64_t)a.src.sample_offset + k0 + amin;
    out.count = n;
  }
}

template <int NU>
__global__ void __launch_bounds__(kReduceThreads) partials_kernel(const ReduceArgs a) { partials_body<NU, kReduceThreads>(a); }

struct CombineArgs {
  const tampc_partial* partials;
  int n_partials;
  int group;             // partials per CTA: CTA b merges [b*group, min((b+1)*group, n_partials))
  int solo;              // 1: the calling CTA merges partials [0, n_partials) whatever its blockIdx (fused tail)
  const SamplerD* sampler;
  const double* U;       // nominal BEFORE the shift
  int T, nu;
  int finalize;          // 1 (single CTA): write U_out / action / stats; 0: write merged tampc_partial b to partial_out[b]
  double* U_out;         // (T, nu) updated nominal sequence
  double* action_out;    // (nu,) = U_out[0]
  double* stats_out;     // beta, eta, argmin
  tampc_partial* partial_out;
  unsigned long long* host_ll;  // optional mapped pinned host memory: action[kMaxNu], stats[3] published straight to the
  unsigned long long seq;       // host, every word tagged with this sequence number (ll_publish: no flag, no fence)
};

constexpr int kCombineThreads = 512;  // also the largest group; combine_body deals partials over 4 x 128 threads
static_assert(kCombineThreads == 512 && TAMPC_MAX_HORIZON * TAMPC_MAX_NU <= 512, "combine_body's thread mapping");

__device__ __forceinline__ void combine_body(const CombineArgs& a) {
  __shared__ double s_scale[kCombineThreads];
  __shared__ double s_wb[32];
  __shared__ long long s_wi[32];
  __shared__ double s_part[4][TAMPC_MAX_HORIZON * kMaxNu];
  const int tid = threadIdx.x;
  const int cta = a.solo ? 0 : (int)blockIdx.x;
  const int p0 = cta * a.group;
  const int n = min(a.group, a.n_partials - p0);
  const tampc_partial* part = a.partials + p0;
  const double li = a.sampler->lambda_inv;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  // ---- min / argmin over the group (lowest global sample index wins ties)
  const double my_beta = tid < n ? part[tid].beta : inf;
  const double my_eta = tid < n ? part[tid].eta : 0.0;
  double beta = my_beta;
  long long arg = tid < n ? part[tid].argmin : 0x7fffffffffffffffll;
  block_min_arg(beta, arg, s_wb, s_wi, tid, kCombineThreads);
  const long long cnt = (long long)block_sum(tid < n ? (double)part[tid].count : 0.0, s_wb, tid, kCombineThreads);
  // ---- log-sum-exp rescale of every partial to the common minimum, and the normaliser
  const double sc = tid < n ? exp(-li * (my_beta - beta)) : 0.0;
  s_scale[tid] = sc;
  const double eta = block_sum(sc * my_eta, s_wb, tid, kCombineThreads);  // (its barriers also publish s_scale)
  // ---- weighted sums.  Partial records are 4 KB apart, so each load is a separate L2 round trip: the partials are
  // dealt over four thread groups (group q takes p = q, q+4, ...), eight loads in flight per thread, and the four
  // group sums are added in group order.  Groups whose scales all underflowed are skipped.
  const int nw = a.T * a.nu;
  {
    const int q = tid >> 7, i = tid & 127;  // kCombineThreads = 4 x 128
    for (int i0 = i; i0 < nw; i0 += 128) {
      double sum = 0.0;
      for (int b0 = q; b0 < n; b0 += 32) {
        double scq[8], v[8];
        bool any = false;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int p = b0 + 4 * e;
          scq[e] = p < n ? s_scale[p] : 0.0;
          any |= scq[e] != 0.0;
        }
        if (!any) continue;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int p = b0 + 4 * e;
          v[e] = (p < n && scq[e] != 0.0) ? part[p].wsum[i0] : 0.0;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) sum = __fma_rn(scq[e], v[e], sum);
      }
      s_part[q][i0] = sum;
    }
  }
  __syncthreads();
  for (int i = tid; i < nw; i += kCombineThreads) {
    const double sum = ((s_part[0][i] + s_part[1][i]) + s_part[2][i]) + s_part[3][i];
    if (a.finalize) {
      const int t = i / a.nu, d = i % a.nu, ts = t + 1;
      const double base = ts < a.T ? a.U[ts * a.nu + d] : a.sampler->u_init[d];
      const double v = base + sum / eta;
      a.U_out[i] = v;
      if (t == 0) {
        a.action_out[d] = v;
        if (a.host_ll != nullptr) ll_publish(a.host_ll, d, v, a.seq);  // zero-copy: straight into mapped pinned host memory
      }
    } else {
      a.partial_out[cta].wsum[i] = sum;
    }
  }
  if (tid == 0) {
    if (a.finalize) {
      a.stats_out[0] = beta;
      a.stats_out[1] = eta;
      a.stats_out[2] = (double)arg;
      if (a.host_ll != nullptr) {
        ll_publish(a.host_ll, kMaxNu, beta, a.seq);
        ll_publish(a.host_ll, kMaxNu + 1, eta, a.seq);
        ll_publish(a.host_ll, kMaxNu + 2, (double)arg, a.seq);
      }
    } else {
      tampc_partial& o = a.partial_out[cta];
      o.beta = beta;
      o.eta = eta;
      o.argmin = arg;
      o.count = cnt;
    }
  }
}

__global__ void __launch_bounds__(kCombineThreads) combine_kernel(const CombineArgs a) { combine_body(a); }

// Both stages in one launch (n_chunks <= kCombineThreads): every CTA writes its partial; the last one to finish
// (device-scope ticket) merges them.  Saves a launch and the merge kernel's cold start on the latency path of a command.
template <int NU>
__global__ void __launch_bounds__(kCombineThreads) partials_combine_kernel(const ReduceArgs r, const CombineArgs c, unsigned int* ticket) {
  __shared__ int s_last;
  partials_body<NU, kCombineThreads>(r);
  __threadfence();  // this CTA's partial is visible device-wide before its ticket is
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = t == gridDim.x - 1;
    if (s_last) *ticket = 0;  // ready for the next command (stream order)
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  combine_body(c);
}

// ---- multi-GPU exchange over peer memory ---------------------------------------------------------------------
// The only inter-GPU step of a K-sharded command is ~1 KB per rank.  Instead of an NCCL all-gather
// (a separate launch plus host-side collective bookkeeping: ~45 us measured at 2 ranks) the merge kernel itself does
// the exchange: every rank stores its merged partial straight into area[rank] of every peer's exchange buffer (P2P
// stores over NVLink to CUDA-IPC mapped memory) with the command's sequence number inside every 8-byte word
// (ll_publish), polls the world's words in its own buffer until they carry that number — no flag, no system-scope
// fence on either side — and then runs the same deterministic combine as the single-GPU path, so every rank
// finishes with a bitwise-identical U.  Buffers are double-buffered on the sequence parity: a rank can run at most
// one command ahead of a peer (it needs the peer's words of command n to finish n), never two.
constexpr int kMaxRanks = 8;
constexpr int kPartialWords = 4 + TAMPC_MAX_HORIZON * TAMPC_MAX_NU;  // a tampc_partial as 64-bit values: beta, eta, argmin, count, wsum
static_assert(sizeof(tampc_partial) == sizeof(unsigned long long) * kPartialWords, "tampc_partial is kPartialWords 64-bit values");
struct ExchangeArgs {
  CombineArgs c;                         // finalize = 1; partials / n_partials are filled in by the kernel
  const tampc_partial* mine;             // this rank's merged partial (device-local)
  unsigned long long* peer_ll[kMaxRanks];  // this parity's area in every rank's buffer: [source rank][kPartialWords][2 words]
  tampc_partial* world;                  // device-local scratch: the world's partials, decoded, in rank order
  int rank, world_size;
  unsigned long long seq;
  long long timeout_ns;
  int* error_out;                        // device int: set to 1 if a peer never arrived
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void exchange_body(const ExchangeArgs& e) {
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  const int nv = 4 + e.c.T * e.c.nu;  // values of a partial in use
  if (tid == 0) s_fail = 0;
  __syncthreads();
  // 1. my partial -> area[rank] on every rank
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(e.mine);
  for (int i = tid; i < nv; i += kCombineThreads) {
    const unsigned long long b = mine[i];
    for (int p = 0; p < e.world_size; ++p) ll_publish_bits(e.peer_ll[p] + (size_t)e.rank * kPartialWords * 2, i, b, e.seq);
  }
  // 2. the world's partials out of my own buffer, as they arrive (bounded wait)
  const unsigned long long t0 = globaltimer_ns();
  for (int it = tid; it < e.world_size * nv; it += kCombineThreads) {
    const int p = it / nv, i = it - p * nv;
    const unsigned long long* src = e.peer_ll[e.rank] + (size_t)p * kPartialWords * 2;
    unsigned long long b = 0;
    int spins = 0;
    while (!ll_try_read(src, i, e.seq, &b)) {
      if (++spins > 64) {
        if ((long long)(globaltimer_ns() - t0) > e.timeout_ns) { s_fail = 1; break; }
        __nanosleep(50);
      }
    }
    reinterpret_cast<unsigned long long*>(e.world + p)[i] = b;
  }
  __syncthreads();  // (also makes the decoded partials visible to the CTA's other threads)
  if (s_fail) {
    if (tid == 0) {
      *e.error_out = 1;
      if (e.c.host_ll != nullptr) ll_publish(e.c.host_ll, kMaxNu + 3, -1.0, e.c.seq);  // wake the host: its error slot
    }
    return;
  }
  // 4. the same merge as on one GPU, over the world's partials in rank order
  CombineArgs c = e.c;
  c.partials = e.world;
  c.n_partials = e.world_size;
  c.group = e.world_size;
  c.solo = 1;
  combine_body(c);
}

__global__ void __launch_bounds__(kCombineThreads) exchange_combine_kernel(const ExchangeArgs e) { exchange_body(e); }

// K-sharded latency path in one launch: chunk partials -> (last CTA) this rank's merged partial -> peer exchange ->
// world merge.  `c_rank` must write its merged partial to e.mine.
template <int NU>
__global__ void __launch_bounds__(kCombineThreads) partials_exchange_kernel(const ReduceArgs r, const CombineArgs c_rank, const ExchangeArgs e,
                                                                            unsigned int* ticket) {
  __shared__ int s_last;
  partials_body<NU, kCombineThreads>(r);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = t == gridDim.x - 1;
    if (s_last) *ticket = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  combine_body(c_rank);
  __threadfence();  // the rank partial written by every thread is visible before any thread forwards it
  __syncthreads();
  exchange_body(e);
}

// ExperimentalMPPI.perturbed_action (K, T, nu), regenerated on request (tampc/env/pybullet_sim.py:127 reads it)
struct PerturbedArgs {
  const SamplerD* sampler;
  const double* U;  // nominal the rollout used, BEFORE the shift
  SampleSrc src;
  int K_local;
  double* out;
  int post_clamp_noise;  // 1: write perturbed - U (controller.py:394) instead of the perturbed action
};

template <int NU>
__global__ void perturbed_kernel(const PerturbedArgs a) {
  __shared__ SamplerS<double> sp;
  stage_sampler<double>(sp, *a.sampler, threadIdx.x, blockDim.x);
  __syncthreads();
  const int T = a.src.T;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)a.K_local * T) return;
  const int k = (int)(i / T), t = (int)(i % T), ts = t + 1;
  double Ut[NU], u[NU], npost[NU];
#pragma unroll
  for (int d = 0; d < NU; ++d) Ut[d] = ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d];
  sample_action<double, NU>(a.src, sp, Ut, k, t, u, npost);
#pragma unroll
  for (int d = 0; d < NU; ++d) a.out[i * NU + d] = a.post_clamp_noise ? npost[d] : u[d];
}

// ---- one-step samplers of the APF baselines (online_controller.py:583-590, 657-664): argmin over the candidates' costs ----
// Single CTA (n <= a few thousand candidates): lowest index wins ties, like torch.argmin.  out[0] = index, out[1] = cost,
// out[2 .. 2+nx) = the next state of the winner (row `index` of states (n, 1, nx)).
__global__ void __launch_bounds__(1024) argmin_select_kernel(const double* __restrict__ cost, const double* __restrict__ states, int n, int nx,
                                                             double* __restrict__ out) {
  __shared__ double s_b[32];
  __shared__ long long s_i[32];
  const int tid = threadIdx.x;
  double b = __longlong_as_double(0x7ff0000000000000ll);
  long long bi = 0x7fffffffffffffffll;
  for (int i = tid; i < n; i += blockDim.x) {
    const double c = cost[i];
    if (c < b || (c == b && i < bi)) { b = c; bi = i; }
  }
  block_min_arg(b, bi, s_b, s_i, tid, blockDim.x);
  if (tid == 0) { out[0] = (double)bi; out[1] = b; }
  if (tid < nx && bi < n) out[2 + tid] = states[(size_t)bi * nx + tid];
}

// goal_cost(x) and the raw trap-set cost with c_u = 1 (U=None) for n states, float64
// (OnlineMPPI.normalize_trapset_cost_to_state, online_controller.py:517-530; cost.py:11-18,154-189).
__global__ void state_costs_kernel(const tampc_cost_desc* __restrict__ c, const double* __restrict__ states, int n, double* __restrict__ goal_out,
                                   double* __restrict__ trap_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int nx = c->nx;
  double x[kMaxNx], d[kMaxNx];
  for (int k = 0; k < nx; ++k) x[k] = states[(size_t)i * nx + k];
  auto diff = [&](const double* g) {
    for (int k = 0; k < nx; ++k) {
      double v = x[k] - g[k];
      if ((c->ang_mask >> k) & 1u) v = angular_wrap_once(v);
      d[k] = v;
    }
  };
  diff(c->goal);
  double q = 0.0;
  for (int j = 0; j < nx; ++j) {
    double col = 0.0;
    for (int k = 0; k < nx; ++k) col = __fma_rn(d[k], c->Q[k * kMaxNx + j], col);
    q = __fma_rn(col, d[j], q);
  }
  goal_out[i] = q;
  double t = 0.0;
  for (int p = 0; p < c->n_traps; ++p) {
    diff(c->trap_x[p]);
    double s = 0.0;
    for (int k = 0; k < nx; ++k) { const double v = d[k] * c->dist_scale[k]; s = __fma_rn(v, v, s); }
    t += fmax(fmin(1.0 / s, kTrapMaxCost), 0.0);
  }
  trap_out[i] = t;
}

}  // namespace tampc
