// generic_rollout.cuh — the fused K x T rollout for networks of ANY layer widths and depths.
//
// The reference builds its networks from configuration (arm_pytorch_utilities `make_sequential_network(h_units=...)`
// as called from tampc/dynamics/model.py:147-170 (util.py:560) and tampc/transform/invariant.py:699-965): the committed
// checkpoints are 16-32 / 32-32 wide, but nothing stops a user from training other shapes.  The tuned kernels
// (rollout_kernel.cuh, mma_rollout.cuh, tc_rollout.cuh) are compiled per shape; this one reads the shape at run time so
// that every `Sequential(Linear, LeakyReLU, ..., Linear)` of 1..TAMPC_MAX_LAYERS Linear layers up to kGenMaxWidth wide
// loads — plain MLP dynamics and the RexExtract triple alike, in float64, float32 and the mixed types.
//
// One thread per sample, like rollout_kernel.cuh, and the same arithmetic in the same order (bias first, inputs in
// ascending order, fused multiply-add; activation max(a, slope a)): for a shape that also has a tuned FMA kernel the
// costs are bitwise identical (tests/test_gpu_generic.py).  What differs is where things live: the layer widths come
// from the kernel parameters (constant bank), the weights are read with warp-uniform 16-byte loads (every lane of a
// warp reads the same address) from shared memory when all networks fit beside the cost program (kGenStageLimit) and
// straight from global memory otherwise, and a thread's activations sit in two ping-pong arrays in local memory
// (L1-resident: 2 x kGenMaxWidth elements).  It is the compatibility path, several times slower than a tuned variant
// of the same shape, and still a GPU kernel: nothing falls back to the host.
#pragma once
#include <type_traits>
#include "rollout_kernel.cuh"

namespace tampc {

// weights and arithmetic of the model's own Linear stacks
template <typename NT>
__device__ __forceinline__ void gen_linear(const NT* w, int IN, int OUT, const NT* __restrict__ in, NT* __restrict__ out, bool act, NT slope) {
  constexpr int V = sizeof(NT) == 4 ? 4 : 2;   // elements per 16-byte load
  constexpr int NV = sizeof(NT) == 4 ? 2 : 4;  // loads per block of outputs: eight independent accumulators either way
  using Vec = typename std::conditional<sizeof(NT) == 4, float4, double2>::type;
  const int OP = pad_out<NT>(OUT);             // row length of the packed block (layer_elems): a multiple of V
#pragma unroll 1
  for (int o0 = 0; o0 < OP; o0 += NV * V) {
    const int nv = min(NV, (OP - o0) / V);     // vectors of this block inside the row (warp-uniform)
    NT acc[NV * V];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      Vec b = *reinterpret_cast<const Vec*>(w + o0 + (c < nv ? c : 0) * V);
      const NT* p = reinterpret_cast<const NT*>(&b);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[c * V + j] = c < nv ? p[j] : (NT)0;
    }
    const NT* wr = w + (size_t)OP + o0;
#pragma unroll 4
    for (int i = 0; i < IN; ++i, wr += OP) {
      const NT x = in[i];
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        if (c < nv) {
          const Vec wv = *reinterpret_cast<const Vec*>(wr + c * V);
          const NT* p = reinterpret_cast<const NT*>(&wv);
#pragma unroll
          for (int j = 0; j < V; ++j) acc[c * V + j] = t_fma(x, p[j], acc[c * V + j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NV * V; ++j) {
      if (o0 + j < OUT) {
        NT a = acc[j];
        if (act) a = t_max(a, slope * a);
        out[o0 + j] = a;
      }
    }
  }
}

// Sequential(Linear, LeakyReLU, ..., Linear): input in a[], returns the buffer that holds the output
template <typename NT>
__device__ __forceinline__ NT* gen_mlp(const GenNet& n, const NT* w, NT* a, NT* b, NT slope) {
#pragma unroll 1
  for (int l = 0; l < n.n; ++l) {
    gen_linear<NT>(w + n.off[l], n.dims[l], n.dims[l + 1], a, b, l + 1 < n.n, slope);
    NT* t = a;
    a = b;
    b = t;
  }
  return a;
}

// one dynamics step (no guard / wrap: the caller applies them, like Model*::predict of device_common.cuh)
template <typename NT, typename ST, int NX, int NU>
__device__ __forceinline__ void gen_predict(const GenDims& g, const NT* w, NT slope, ST (&x)[NX], const ST (&u)[NU], NT* b0, NT* b1) {
#pragma unroll
  for (int i = 0; i < NX; ++i) b0[i] = (NT)x[i];
#pragma unroll
  for (int i = 0; i < NU; ++i) b0[NX + i] = (NT)u[i];
  if (g.kind == TAMPC_DYN_MLP) {
    const NT* dx = gen_mlp<NT>(g.net, w, b0, b1, slope);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] += (ST)dx[i];
    return;
  }
  // RexExtract (invariant.py:792-795,926-931) + latent dynamics (model.py:300-301), folded like ModelRex
  NT v[kMaxNx];
  {
    NT* z = gen_mlp<NT>(g.enc, w, b0, b1, slope);
    NT* other = z == b0 ? b1 : b0;
    const NT* vv = gen_mlp<NT>(g.net, w, z, other, slope);
#pragma unroll
    for (int j = 0; j < kMaxNx; ++j) v[j] = j < g.nv ? vv[j] : (NT)0;
  }
#pragma unroll
  for (int i = 0; i < NX; ++i) b0[i] = (NT)x[i];
  const NT* C = gen_mlp<NT>(g.dec, w, b0, b1, slope);
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    NT dx = w[g.off_d0 + i];
    for (int j = 0; j < g.nv; ++j) dx = t_fma(C[i * g.nv + j], v[j], dx);
    x[i] += (ST)dx;
  }
}

// shared memory: CostS | SamplerS | shifted nominal sequence | the networks, if they fit (else they stay in global memory)
constexpr int kGenStageLimit = 150 * 1024;  // bytes of shared memory per CTA up to which the weights are staged
template <typename ST, typename CT>
struct GenericSmem {
  static constexpr size_t kCost = 0;
  static constexpr size_t kSampler = align16<int>(sizeof(CostS<ST, CT>));
  static constexpr size_t kU = kSampler + align16<int>(sizeof(SamplerS<ST>));
  static __host__ __device__ constexpr size_t bytes(int T) { return kU + align16<int>(sizeof(ST) * (size_t)T * kMaxNu); }
};

template <typename NT, typename ST, typename CT, int NX, int NU>
__global__ void __launch_bounds__(128) rollout_generic_kernel(const RolloutArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using L = GenericSmem<ST, CT>;
  CostS<ST, CT>& cs = *reinterpret_cast<CostS<ST, CT>*>(smem_raw + L::kCost);
  SamplerS<ST>& sp = *reinterpret_cast<SamplerS<ST>*>(smem_raw + L::kSampler);
  ST* Ush = reinterpret_cast<ST*>(smem_raw + L::kU);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int T = a.src.T;
  const NT* w = reinterpret_cast<const NT*>(a.image);
  {
    grid_dep_launch();
    // the image is  networks | CostS | SamplerS  (mppi_api.cu install_model): the last two go to shared memory ...
    copy_image_async(smem_raw, reinterpret_cast<const unsigned char*>(a.image) + a.gen.off_cost, (int)L::kU, tid, nthr);
    if (a.gen.stage_w) {  // ... and the networks behind them when the launcher found room
      unsigned char* ws = smem_raw + L::bytes(T);
      copy_image_async(ws, a.image, a.gen.off_cost, tid, nthr);
      w = reinterpret_cast<const NT*>(ws);
    }
    for (int i = tid; i < T * NU; i += nthr) {
      const int t = i / NU, d = i % NU, ts = t + a.shift;
      Ush[t * kMaxNu + d] = (ST)(ts < T ? a.U[ts * NU + d] : a.sampler->u_init[d]);
    }
    copy_image_wait();
  }
  __syncthreads();
  const int g = blockIdx.x * nthr + tid;
  const bool live = g < a.K_local;
  const int k = live ? g : a.K_local - 1;  // idle lanes recompute the last sample, never write
  ST x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = (ST)a.state[i];
  ST cost = (ST)0, pert = (ST)0;
  const NT slope = (NT)a.slope;
  const ST bad2 = (ST)(a.bad_norm * a.bad_norm);
  NT buf0[kGenMaxWidth], buf1[kGenMaxWidth];
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    ST u[NU], npost[NU];
    pert += sample_action<ST, NU>(a.src, sp, Ush + t * kMaxNu, k, t, u, npost);
    bool bad = false;
    if (a.has_guard) {  // OnlineMPC._apply_dynamics bad-state hack (online_controller.py:62-66)
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
    gen_predict<NT, ST, NX, NU>(a.gen, w, slope, x, u, buf0, buf1);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      if ((a.wrap_mask >> i) & 1u) x[i] = angle_normalize<ST>(x[i]);  // constrain_state
      if (bad) x[i] = (ST)0;                                          // next_state[bad] = state[bad] (= 0)
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

template <typename NT, typename ST, typename CT, int NX, int NU>
cudaError_t launch_rollout_generic(const RolloutArgs& a, int block, cudaStream_t stream) {
  auto kern = rollout_generic_kernel<NT, ST, CT, NX, NU>;
  size_t smem = GenericSmem<ST, CT>::bytes(a.src.T);
  RolloutArgs b = a;
  b.gen.stage_w = smem + (size_t)a.gen.off_cost <= (size_t)kGenStageLimit ? 1 : 0;
  if (b.gen.stage_w) {
    smem += (size_t)a.gen.off_cost;
    block = 128;  // four warps share one copy of the networks
  }
  static size_t configured = 0;  // per instantiation
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int grid = (a.K_local + block - 1) / block;
  kern<<<grid, block, smem, stream>>>(b);
  return cudaGetLastError();
}

// precision class: 0 = float64 everywhere, 1 = float32 networks with float64 state (MIXED and the tensor precisions),
// 2 = float32 everywhere.  nullptr: (nx, nu) pair not instantiated.
rollout_launch_fn generic_launcher(int nx, int nu, int types);

}  // namespace tampc
