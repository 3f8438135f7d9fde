// rollout_kernel.cuh — the fused K x T rollout kernel (one launch per MPPI command).
//
// Replaces ExperimentalMPPI._compute_total_cost_batch + _compute_rollout_costs
// (tampc/controller/controller.py:340-402): every thread owns S samples, keeps their state in registers for the
// whole horizon, and writes one float64 cost per sample; no trajectory is written unless states_out is given
// (the reference stacks M*K*T*nx float64 states, controller.py:367-373).
#pragma once
#include "device_common.cuh"

namespace tampc {

// Shape and image layout of the tcgen05 plain-MLP kernel (tc_rollout.cuh), filled in by mppi_api.cu pack_tc_mlp.
struct TcMlpDims {
  int nin;        // nx + nu
  int k0_steps;   // K = 8 steps of the input row (1 or 2)
  int h1, h2;     // hidden widths, multiples of nc, <= 256
  int nc;         // width of a layer-1 output chunk = K-chunk of layer 2 (32 or 16)
  int n_chunks;   // h1 / nc
  int stages;     // ring stages, each holding one K = 8 step of W2 (hi slab + lo slab)
  int resident;   // h1 / 8 <= stages: W2 is loaded once and stays
  int n_issuers;  // threads issuing layer 2 (1, or 2: each half of the output columns)
  int off_w1;     // image: per chunk [k0 hi slabs | k0 lo slabs | bias slab], slab = (nc/8) * 256 bytes
  int off_b2;     // image: bias slab of layer 2, (h2/8) * 256 bytes
  int off_w3;     // image: float [h2][2 * ceil(nx/2)] (output pairs for packed FFMA2)
  int off_b3;     // image: float [8]
  int off_cost, off_sampler;  // image: CostS<float,float>, SamplerS<float>
  int w1_chunk_bytes, w2_stage_bytes;
};

// runtime-shaped networks (generic_rollout.cuh): Sequential(Linear, LeakyReLU, ..., Linear) of n Linear layers
// dims[0] -> ... -> dims[n]; off[l] = element offset of layer l's packed block (layer_elems layout) in the image
constexpr int kGenMaxWidth = 256;
struct GenNet {
  int n;
  int dims[TAMPC_MAX_LAYERS + 1];
  int off[TAMPC_MAX_LAYERS];
};
struct GenDims {
  int kind;      // TAMPC_DYN_MLP: `net` is (x,u) -> dx;  TAMPC_DYN_REX: enc (x,u) -> z, net z -> v, dec x -> C (nx*nv)
  int nv;        // REX: latent output width
  int off_d0;    // REX: element offset of d0 (nx values)
  int off_cost;  // byte offset of CostS in the image (SamplerS and nothing else follow)
  int stage_w;   // set by the launcher: the weights ([0, off_cost) of the image) are staged in shared memory
  GenNet enc, net, dec;
};

struct RolloutArgs {
  const void* image;    // shared-memory image: packed networks | CostS | SamplerS in the kernel's types
  int image_bytes;
  double slope;
  const tampc_cost_desc* cost;
  const SamplerD* sampler;
  const double* U;      // nominal sequence (T, nu) BEFORE the shift
  int shift;            // 1: step t uses U[t+1] (u_init at the tail) = MPPI.command's roll; 0: U[t]
  double state[kMaxNx];  // start state shared by all samples, passed by value (no H2D copy, no dependent load)
  SampleSrc src;
  int K_local;
  double* cost_out;     // (K_local,)
  double* states_out;   // optional (K_local, T, nx)
  uint32_t wrap_mask;
  int has_guard;
  double bad_norm;
  // fused softmax update (fused_reduce.cuh; small-K tensor-core kernels only).  0: off, separate partials / combine
  // kernels follow; 1: the last warp finalises U / action; 2: the last warp writes this rank's merged partial.
  int fuse;
  struct PartialHdr* phdr;  // per-warp partial headers (one slot per warp of the grid) ...
  double* pwsum;            // ... and weighted sums, stride T*nu
  unsigned int* ticket;     // zero-initialised; reset by the last warp
  double* U_out;
  double* action_out;
  double* stats_out;
  unsigned long long* host_ll;  // mapped pinned host memory (action, stats; ll_publish) or null
  unsigned long long seq;
  tampc_partial* partial_out;
  // online GP residual (gp_rollout.cuh): shared-memory block, M rollout replicas per sample, variance penalty,
  // optional injected standard normals (K_local, M, T, n_out)
  const void* gp;
  int gp_bytes;
  int M;
  double var_cost, var_discount;
  const double* gp_eps;
  // runtime-shaped networks (generic_rollout.cuh)
  GenDims gen;
  // tcgen05 plain-MLP kernel: W2 in K-chunk-major slab order in global memory (streamed by TMA), dims
  const void* tc_w2;
  TcMlpDims tc;
  long long* tc_prof;  // development: per-role clock64() stamps of CTA 0 at step 2 (TAMPC_TC_PROF=1), else null
};

template <typename T> __host__ __device__ constexpr size_t align16(size_t n) { return (n + 15) & ~size_t(15); }

template <class Model, typename NT, typename ST, typename CT>
struct RolloutSmem {
  static constexpr size_t kWeights = 0;
  static constexpr size_t kCost = align16<int>(sizeof(NT) * Model::kElems);
  static constexpr size_t kSampler = kCost + align16<int>(sizeof(CostS<ST, CT>));
  static constexpr size_t kU = kSampler + align16<int>(sizeof(SamplerS<ST>));
  static __host__ __device__ constexpr size_t bytes(int T) { return kU + align16<int>(sizeof(ST) * (size_t)T * kMaxNu); }
};

template <class Model, typename NT, typename ST, typename CT, int S>
__global__ void __launch_bounds__(128) rollout_kernel(const RolloutArgs a) {
  constexpr int NX = Model::NX, NU = Model::NU;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = RolloutSmem<Model, NT, ST, CT>;
  NT* w = reinterpret_cast<NT*>(smem_raw + L::kWeights);
  CostS<ST, CT>& cs = *reinterpret_cast<CostS<ST, CT>*>(smem_raw + L::kCost);
  SamplerS<ST>& sp = *reinterpret_cast<SamplerS<ST>*>(smem_raw + L::kSampler);
  ST* Ush = reinterpret_cast<ST*>(smem_raw + L::kU);  // [T][kMaxNu] shifted nominal sequence

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int T = a.src.T;
  {  // ---- stage the image (weights, cost program, sampler constants) and the (shifted) nominal sequence
    copy_image_async(smem_raw, a.image, (int)L::kU, tid, nthr);
    for (int i = tid; i < T * NU; i += nthr) {
      int t = i / NU, d = i % NU;
      int ts = t + a.shift;
      Ush[t * kMaxNu + d] = (ST)(ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d]);
    }
    copy_image_wait();
  }
  __syncthreads();

  const int g = blockIdx.x * nthr + tid;
  int k[S];
  bool live[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    int kk = g * S + s;
    live[s] = kk < a.K_local;
    k[s] = live[s] ? kk : a.K_local - 1;  // idle lanes recompute the last sample, never write
  }

  ST x[S][NX];
  ST cost[S], pert[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
#pragma unroll
    for (int i = 0; i < NX; ++i) x[s][i] = (ST)a.state[i];
    cost[s] = (ST)0;
    pert[s] = (ST)0;
  }
  const NT slope = (NT)a.slope;
  const ST bad2 = (ST)(a.bad_norm * a.bad_norm);

#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    ST u[S][NU];
    bool bad[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      ST npost[NU];
      pert[s] += sample_action<ST, NU>(a.src, sp, Ush + t * kMaxNu, k[s], t, u[s], npost);
      bad[s] = false;
      if (a.has_guard) {  // OnlineMPC._apply_dynamics bad-state hack (online_controller.py:62-66)
        ST n2 = (ST)0;
#pragma unroll
        for (int i = 0; i < NX; ++i) n2 = t_fma(x[s][i], x[s][i], n2);
        bad[s] = n2 > bad2;
        if (bad[s]) {
#pragma unroll
          for (int i = 0; i < NX; ++i) x[s][i] = (ST)0;
#pragma unroll
          for (int i = 0; i < NU; ++i) u[s][i] = (ST)0;
        }
      }
    }
    Model::template predict<ST, S>(w, slope, x, u);
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        if ((a.wrap_mask >> i) & 1u) x[s][i] = angle_normalize<ST>(x[s][i]);  // constrain_state
        if (bad[s]) x[s][i] = (ST)0;                                           // next_state[bad] = state[bad] (= 0)
      }
      cost[s] += (ST)running_cost<ST, CT, NX, NU>(cs, x[s], u[s]);
      if (a.states_out != nullptr && live[s]) {
        double* so = a.states_out + ((size_t)k[s] * T + t) * NX;
#pragma unroll
        for (int i = 0; i < NX; ++i) so[i] = (double)x[s][i];
      }
    }
  }
#pragma unroll
  for (int s = 0; s < S; ++s) {
    ST total = cost[s] + (ST)terminal_cost<ST, CT, NX>(cs, x[s]);
    total += pert[s];
    if (live[s]) a.cost_out[k[s]] = (double)total;
  }
}

// host-callable launcher signature shared by all variants
typedef cudaError_t (*rollout_launch_fn)(const RolloutArgs& a, int block, cudaStream_t stream);

template <class Model, typename NT, typename ST, typename CT, int S>
cudaError_t launch_rollout(const RolloutArgs& a, int block, cudaStream_t stream) {
  auto kern = rollout_kernel<Model, NT, ST, CT, S>;
  const size_t smem = RolloutSmem<Model, NT, ST, CT>::bytes(a.src.T);
  static size_t configured = 0;  // per instantiation
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int per_block = block * S;
  const int grid = (a.K_local + per_block - 1) / per_block;
  kern<<<grid, block, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace tampc
