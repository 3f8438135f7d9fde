// mppi_api.cu — C ABI of libtampc_mppi.so (include/tampc_mppi.h): handle, model packing, command sequencing.
//
// Host-side shell of pytorch_mppi `MPPI` as used by tampc (tampc/controller/controller.py:316-435):
//   command()        -> rollout_kernel + partials_kernel + combine_kernel on one stream, one pinned D2H of U[0]
//   set_model()      -> folds the scaler chain of tampc/util.py:408-411 into the network layers and lays the
//                       weights out for broadcast LDS.128 (device_common.cuh `linear`)
//   set_cost()       -> cost program by value (mutated between calls: online_controller.py:400-402,416-418)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <string>
#include <vector>

#include "reduce_kernels.cuh"
#include "variants.cuh"
#include "tc_rollout.cuh"
#include "generic_rollout.cuh"

using namespace tampc;

namespace {

thread_local std::string g_create_error;

#define CUDA_TRY(h, expr)                                                                            \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                                \
      return TAMPC_ERR_CUDA;                                                                         \
    }                                                                                                \
  } while (0)

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

}  // namespace

struct tampc_mppi {
  tampc_mppi_config cfg{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;

  const ModelVariant* variant = nullptr;
  tampc_model_desc model{};
  bool has_model = false, has_cost = false, has_nominal = false;
  int T = 0;
  int n_sm = 148;
  uint64_t call = 0;
  long long launches = 0;

  SamplerD* d_sampler = nullptr;
  tampc_cost_desc* d_cost = nullptr;
  double* d_U[2] = {nullptr, nullptr};
  int cur = 0;           // d_U[cur] = current nominal sequence
  double* d_state = nullptr;
  double* d_costs = nullptr;
  double* d_noise = nullptr;    // injected noise / given actions staging (K_local, Tmax, nu)
  double* d_states = nullptr;   // optional trajectories (K_local, Tmax, nx)
  double* d_scratch = nullptr;  // perturbed-action readback
  double* d_predict = nullptr;  // action + cost scratch of tampc_mppi_predict
  double* d_select = nullptr;   // index, cost, next state of tampc_mppi_select_action
  tampc_partial* d_partials = nullptr;
  int n_partial_slots = 0;
  // peer-memory exchange (multi-GPU): [parity][source rank][kPartialWords values][2 tagged words] (reduce_kernels.cuh)
  int comm_rank = 0, comm_world = 0;
  bool comm_connected = false;
  bool comm_broken = false;  // a peer exchange timed out: U / sequence numbers of the ranks may disagree until a reconnect
  unsigned char* comm_local = nullptr;
  unsigned char* comm_peer[kMaxRanks] = {nullptr};
  int* d_comm_error = nullptr;
  tampc_partial* d_world_partials = nullptr;  // the world's partials decoded out of the exchange buffer (kMaxRanks)
  unsigned long long comm_seq = 0;
  unsigned int* d_ticket = nullptr;      // fused-reduction ticket counter
  PartialHdr* d_phdr = nullptr;          // fused-reduction per-warp partials (compact)
  double* d_pwsum = nullptr;
  tampc_partial* d_partials2 = nullptr;  // second-level combine scratch
  int n_partials2 = 0;
  tampc_partial* d_rank_partial = nullptr;
  double* d_action = nullptr;
  double* d_stats = nullptr;
  void* d_weights = nullptr;
  size_t weights_bytes = 0;
  int n_weight_elems = 0;
  bool use_mma = false;  // the staged weights are in the mma-path layout
  bool use_tc = false;   // tcgen05 plain-MLP kernel (tc_rollout.cuh): image in its slab layout, W2 streamed from d_tc_w2
  bool use_generic = false;  // runtime-shaped networks (generic_rollout.cuh): FMA layout read from global memory
  GenDims gen{};
  rollout_launch_fn gen_fn = nullptr;
  const char* gen_name = "";
  TcMlpDims tc_dims{};
  rollout_launch_fn tc_fn = nullptr;
  void* d_tc_w2 = nullptr;
  size_t tc_w2_bytes = 0;
  long long* d_tc_prof = nullptr;  // development (TAMPC_TC_PROF=1): clock stamps of the tcgen05 kernel's roles
  int img_cost_off = 0, img_sampler_off = 0, img_sampler_d_off = 0, img_bytes = 0;  // shared-memory image offsets inside d_weights

  double* h_pin = nullptr;  // pinned staging: [0..15] in, [16..31] out
  unsigned long long* h_map = nullptr;  // mapped pinned result block published by the combine kernel (ll_publish): kHostSlots x 2 words
  unsigned long long* d_map = nullptr;  // device alias of h_map
  unsigned long long seq = 0;
  double state_host[kMaxNx] = {0};  // start state of the next rollout (passed to the kernels by value)
  int n_chunks = 0, chunk = 0;

  // last command bookkeeping (for queries)
  int last_mode = kModePhilox;
  uint64_t last_call = 0;
  int last_T = 0;
  int last_U = 0;  // index of the pre-shift nominal the last command used
  bool last_valid = false;
  bool states_valid = false;
  double last_stats[3] = {0, 0, 0};

  // online GP residual (tampc_mppi_set_gp)
  std::vector<double> model_params;  // kept so that the network image can be re-laid-out when the GP toggles
  double net_in_scale[kMaxNx + kMaxNu] = {0}, net_in_off[kMaxNx + kMaxNu] = {0}, net_out_scale[kMaxNx] = {0};  // scalers around `net`
  bool gp_on = false;
  bool img_f64 = false;  // the installed image holds float64 networks (precision F64, or a GP with the replica-variance penalty)
  tampc_gp_desc gp{};
  void* d_gp = nullptr;
  size_t gp_capacity = 0;
  int gp_bytes = 0;
  double* d_gp_eps = nullptr;
  size_t gp_eps_capacity = 0;
  bool gp_eps_armed = false;
  int gp_eps_T = 0;
  uint64_t gp_aux_calls = 0;
  bool tail_exchanged = false;  // the last fused tail also did the peer exchange
};

namespace {

// ------------------------------------------------------------------------------------------------
// model packing (all folding in float64, then converted to the network type)

struct HostLayer {
  int in = 0, out = 0;
  std::vector<double> W;  // (out, in) row-major
  std::vector<double> b;
};

int read_mlp(tampc_mppi* h, const tampc_mlp_desc& d, const double* p, size_t n, std::vector<HostLayer>& out) {
  if (d.n_layers < 1 || d.n_layers > TAMPC_MAX_LAYERS) { h->err = "mlp n_layers out of range"; return TAMPC_ERR_INVALID; }
  out.resize(d.n_layers);
  for (int l = 0; l < d.n_layers; ++l) {
    HostLayer& L = out[l];
    L.in = d.dims[l];
    L.out = d.dims[l + 1];
    if (L.in < 1 || L.out < 1) { h->err = "mlp layer with non-positive width"; return TAMPC_ERR_INVALID; }
    if (d.w_off[l] < 0 || d.b_off[l] < 0 || (size_t)d.w_off[l] + (size_t)L.in * L.out > n || (size_t)d.b_off[l] + L.out > n) {
      h->err = "mlp weight offsets outside the params array";
      return TAMPC_ERR_INVALID;
    }
    L.W.assign(p + d.w_off[l], p + d.w_off[l] + (size_t)L.in * L.out);
    L.b.assign(p + d.b_off[l], p + d.b_off[l] + L.out);
  }
  return TAMPC_OK;
}

int read_affine(tampc_mppi* h, const tampc_affine_desc& a, const double* p, size_t n, int dim, std::vector<double>& scale,
                std::vector<double>& offset) {
  scale.assign(dim, 1.0);
  offset.assign(dim, 0.0);
  if (a.scale_off < 0 && a.offset_off < 0) return TAMPC_OK;
  if (a.scale_off < 0 || a.offset_off < 0 || (size_t)a.scale_off + dim > n || (size_t)a.offset_off + dim > n) {
    h->err = "affine offsets outside the params array";
    return TAMPC_ERR_INVALID;
  }
  for (int i = 0; i < dim; ++i) {
    scale[i] = p[a.scale_off + i];
    offset[i] = p[a.offset_off + i];
    if (scale[i] == 0.0 || !std::isfinite(scale[i])) { h->err = "affine scale must be finite and non-zero"; return TAMPC_ERR_INVALID; }
  }
  return TAMPC_OK;
}

// y = W (x*s + o) + b  ->  W' = W diag(s), b' = b + W o
void fold_input_affine(HostLayer& L, const std::vector<double>& s, const std::vector<double>& o) {
  for (int r = 0; r < L.out; ++r) {
    double acc = L.b[r];
    for (int c = 0; c < L.in; ++c) {
      acc += L.W[(size_t)r * L.in + c] * o[c];
      L.W[(size_t)r * L.in + c] *= s[c];
    }
    L.b[r] = acc;
  }
}
// y' = y*s + o applied after the layer
void fold_output_affine(HostLayer& L, const std::vector<double>& s, const std::vector<double>& o) {
  for (int r = 0; r < L.out; ++r) {
    for (int c = 0; c < L.in; ++c) L.W[(size_t)r * L.in + c] *= s[r];
    L.b[r] = L.b[r] * s[r] + o[r];
  }
}

template <typename NT>
void pack_layer(std::vector<NT>& blob, const HostLayer& L) {
  const int OP = pad_out<NT>(L.out);
  const size_t base = blob.size();
  blob.resize(base + (size_t)(L.in + 1) * OP, (NT)0);
  for (int j = 0; j < L.out; ++j) blob[base + j] = (NT)L.b[j];
  for (int i = 0; i < L.in; ++i)
    for (int j = 0; j < L.out; ++j) blob[base + (size_t)(i + 1) * OP + j] = (NT)L.W[(size_t)j * L.in + i];
}

template <typename NT>
void pack_mlp(std::vector<NT>& blob, const std::vector<HostLayer>& m) {
  for (const HostLayer& L : m) pack_layer<NT>(blob, L);
}

// ---- mma-path layout (mma_rollout.cuh): per layer  bias[out8*8]  then  in8*out8 blocks of 32 lanes x 16 bytes.
// Lane (g = lane/4, t = lane%4) of block (n, o) holds W[8o+g][8n+2t] and W[8o+g][8n+2t+1]: as two doubles (DMMA
// k-steps 2n, 2n+1) or as {hi0, hi1, lo0, lo1} TF32 pieces for the 3xTF32 split.
float tf32_rna(double v) {
  float f = (float)v;
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&f, &u, 4);
  return f;
}

template <typename T>
void append(std::vector<unsigned char>& blob, T v) {
  const unsigned char* p = reinterpret_cast<const unsigned char*>(&v);
  blob.insert(blob.end(), p, p + sizeof(T));
}

void pack_layer_mma(std::vector<unsigned char>& blob, const HostLayer& L, int in8, int out8, bool f64) {
  auto W = [&](int o, int i) { return (o < L.out && i < L.in) ? L.W[(size_t)o * L.in + i] : 0.0; };
  for (int j = 0; j < out8 * 8; ++j) {
    const double b = j < L.out ? L.b[j] : 0.0;
    if (f64) append<double>(blob, b); else append<float>(blob, (float)b);
  }
  for (int n = 0; n < in8; ++n)
    for (int o = 0; o < out8; ++o)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        const double w0 = W(8 * o + g, 8 * n + 2 * t), w1 = W(8 * o + g, 8 * n + 2 * t + 1);
        if (f64) {
          append<double>(blob, w0);
          append<double>(blob, w1);
        } else {
          const float h0 = tf32_rna(w0), h1 = tf32_rna(w1);
          append<float>(blob, h0);
          append<float>(blob, h1);
          append<float>(blob, tf32_rna(w0 - (double)h0));
          append<float>(blob, tf32_rna(w1 - (double)h1));
        }
      }
}

// 3xFP16 layout (mma_rollout.cuh mma_layer_f16): per layer  bias[out8*8] floats  then  ceil(in8/2)*out8 blocks of
// 32 lanes x 16 bytes.  Lane (g,t) of block (k, o) holds, for output 8o+g, the inputs 16k + {2t, 2t+1} (register b0) and
// 16k + {2t+8, 2t+9} (b1) as half2 pairs: {b0 hi, b1 hi, b0 lo', b1 lo'} with hi = fp16(w), lo' = fp16((w - hi) * 2^11).
void pack_layer_f16(std::vector<unsigned char>& blob, const HostLayer& L, int in8, int out8) {
  auto W = [&](int o, int i) { return (o < L.out && i < L.in) ? L.W[(size_t)o * L.in + i] : 0.0; };
  auto split = [](double w, uint16_t& hi, uint16_t& lo) {
    float wf = (float)w;
    if (wf > 65504.f) wf = 65504.f;
    if (wf < -65504.f) wf = -65504.f;
    const __half h = __float2half_rn(wf);
    const __half l = __float2half_rn((float)((w - (double)__half2float(h)) * 2048.0));
    memcpy(&hi, &h, 2);
    memcpy(&lo, &l, 2);
  };
  for (int j = 0; j < out8 * 8; ++j) append<float>(blob, j < L.out ? (float)L.b[j] : 0.f);
  const int in16 = (in8 + 1) / 2;
  for (int k = 0; k < in16; ++k)
    for (int o = 0; o < out8; ++o)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3, n = 8 * o + g;
        const int idx[4] = {16 * k + 2 * t, 16 * k + 2 * t + 1, 16 * k + 2 * t + 8, 16 * k + 2 * t + 9};
        uint16_t hi[4], lo[4];
        for (int e = 0; e < 4; ++e) split(idx[e] < in8 * 8 ? W(n, idx[e]) : 0.0, hi[e], lo[e]);
        for (int e = 0; e < 4; ++e) append<uint16_t>(blob, hi[e]);
        for (int e = 0; e < 4; ++e) append<uint16_t>(blob, lo[e]);
      }
}

// kind: 0 = 3xTF32, 1 = float64 (DMMA), 2 = 3xFP16
void pack_mlp_mma(std::vector<unsigned char>& blob, const std::vector<HostLayer>& m, int out8_last, int kind);
void pack_mlp_mma(std::vector<unsigned char>& blob, const std::vector<HostLayer>& m, int out8_last, bool f64) { pack_mlp_mma(blob, m, out8_last, f64 ? 1 : 0); }

void pack_mlp_mma(std::vector<unsigned char>& blob, const std::vector<HostLayer>& m, int out8_last, int kind) {
  int in8k = 1;
  if (kind == 2) {
    for (size_t l = 0; l < m.size(); ++l) {
      const int out8 = (l + 1 == m.size()) ? out8_last : m[l].out / 8;
      pack_layer_f16(blob, m[l], in8k, out8);
      in8k = out8;
    }
    return;
  }
  const bool f64 = kind == 1;
  int in8 = 1;
  for (size_t l = 0; l < m.size(); ++l) {
    const int out8 = (l + 1 == m.size()) ? out8_last : m[l].out / 8;
    pack_layer_mma(blob, m[l], in8, out8, f64);
    in8 = out8;
  }
}

bool dims_are(const std::vector<HostLayer>& m, std::initializer_list<int> dims) {
  if (m.size() + 1 != dims.size()) return false;
  auto it = dims.begin();
  if (m[0].in != *it) return false;
  ++it;
  for (size_t l = 0; l < m.size(); ++l, ++it)
    if (m[l].out != *it) return false;
  return true;
}

// placeholder variant of the shape-generic tcgen05 kernel (no per-shape kernels; the launcher lives in the handle)
const ModelVariant* variant_tc_generic() {
  static const ModelVariant v = [] {
    ModelVariant m{};
    m.name = "mlp_tcgen05_generic";
    m.kind = TAMPC_DYN_MLP;
    return m;
  }();
  return &v;
}

struct PackedModel {
  const ModelVariant* variant = nullptr;
  std::vector<float> f32;              // FMA-path layouts
  std::vector<double> f64;
  std::vector<unsigned char> mma_f64;  // mma-path layouts
  std::vector<unsigned char> mma_tf32;
  std::vector<unsigned char> mma_f16;
  // tcgen05 plain-MLP kernel
  bool tc = false;
  std::vector<unsigned char> tc_image, tc_w2;
  TcMlpDims tc_dims{};
  rollout_launch_fn tc_fn = nullptr;
  // runtime-shaped networks (generic_rollout.cuh): f32 / f64 above hold the FMA-layout blobs, these their layer tables
  bool generic = false;
  GenDims gen_f32{}, gen_f64{};
};

// ---- tcgen05 layout (tc_rollout.cuh): K-major no-swizzle slabs.  One slab = rows [n0, n0 + N) x one K = 8 step:
// (N/8) groups of 256 bytes = [k half (2)][row in group (8)][k in half (4)] floats — the 8 x 16-byte core matrices the
// shared-memory descriptor walks with LBO = 128, SBO = 256.
template <typename F>
void append_slab(std::vector<unsigned char>& blob, int n0, int N, F value) {
  for (int g = 0; g < N / 8; ++g)
    for (int half = 0; half < 2; ++half)
      for (int r = 0; r < 8; ++r)
        for (int kk = 0; kk < 4; ++kk) append<float>(blob, value(n0 + g * 8 + r, half * 4 + kk));
}

// (nx+nu) -> h1 -> h2 -> nx, scalers already folded into `net`.  Image: W1 chunk by chunk [hi slabs | lo slabs | bias slab],
// layer-2 bias slab, W3 as float [h2][8] and b3 (layer 3 runs on the CUDA cores); W2 separately in K-chunk-major order for TMA.
void pack_tc_mlp(const std::vector<HostLayer>& net, int nx, int nu, PackedModel& out) {
  TcMlpDims d{};
  d.nin = nx + nu;
  d.k0_steps = (d.nin + 7) / 8;
  // hidden widths are padded to multiples of 16 with zero weights and biases: a padded unit is leaky(0) = 0 and feeds zeros
  d.h1 = (net[0].out + 15) & ~15;
  d.h2 = (net[1].out + 15) & ~15;
  d.nc = (d.h1 % 32 == 0 && d.h2 % 32 == 0) ? 32 : 16;
  d.n_chunks = d.h1 / d.nc;
  auto hi = [](double w) { return tf32_rna(w); };
  auto lo = [](double w) { return tf32_rna(w - (double)tf32_rna(w)); };
  auto W = [](const HostLayer& L, int o, int i) { return (o < L.out && i < L.in) ? L.W[(size_t)o * L.in + i] : 0.0; };
  auto Bv = [](const HostLayer& L, int o) { return o < L.out ? L.b[o] : 0.0; };
  std::vector<unsigned char>& img = out.tc_image;
  d.off_w1 = 0;
  for (int c = 0; c < d.n_chunks; ++c) {
    const size_t before = img.size();
    for (int ks = 0; ks < d.k0_steps; ++ks) append_slab(img, c * d.nc, d.nc, [&](int n, int k) { return hi(W(net[0], n, ks * 8 + k)); });
    for (int ks = 0; ks < d.k0_steps; ++ks) append_slab(img, c * d.nc, d.nc, [&](int n, int k) { return lo(W(net[0], n, ks * 8 + k)); });
    append_slab(img, c * d.nc, d.nc, [&](int n, int k) { return k == 0 ? hi(Bv(net[0], n)) : (k == 1 ? lo(Bv(net[0], n)) : 0.f); });
    d.w1_chunk_bytes = (int)(img.size() - before);
  }
  d.off_b2 = (int)img.size();
  append_slab(img, 0, d.h2, [&](int n, int k) { return k == 0 ? hi(Bv(net[1], n)) : (k == 1 ? lo(Bv(net[1], n)) : 0.f); });
  d.off_w3 = (int)img.size();
  for (int k = 0; k < d.h2; ++k)  // [k][ceil(nx/2)] pairs of outputs (packed FFMA2 operands)
    for (int j = 0; j < 2 * ((nx + 1) / 2); ++j) append<float>(img, (float)W(net[2], j, k));
  while (img.size() % 16) append<float>(img, 0.f);
  d.off_b3 = (int)img.size();
  for (int j = 0; j < 8; ++j) append<float>(img, j < nx ? (float)net[2].b[j] : 0.f);
  std::vector<unsigned char>& w2 = out.tc_w2;  // K-step-major: [k-step][hi slab | lo slab], one ring stage each
  for (int q = 0; q < d.h1 / 8; ++q) {
    const size_t before = w2.size();
    append_slab(w2, 0, d.h2, [&](int n, int k) { return hi(W(net[1], n, q * 8 + k)); });
    append_slab(w2, 0, d.h2, [&](int n, int k) { return lo(W(net[1], n, q * 8 + k)); });
    d.w2_stage_bytes = (int)(w2.size() - before);
  }
  out.tc = true;
  out.tc_dims = d;
  out.tc_fn = tc_mlp_launcher(nx, nu);
}

// Does the tcgen05 kernel take this MLP?  TAMPC_TC=0 never, =1 whenever the shape allows; default: shapes without a tuned
// kernel of their own (any make_sequential_network(h_units=...) widths), and the wide shapes at throughput-bound K.
bool tc_mlp_wanted(const tampc_mppi* h, const std::vector<HostLayer>& net, int nx, int nu, bool has_tuned_variant) {
  if (h->cfg.precision != TAMPC_PREC_TF32X3 || net.size() != 3) return false;
  const int h1 = net[0].out, h2 = net[1].out;
  if (h1 > 256 || h2 > 256 || nx > 8 || nu > kMaxNu || !tc_mlp_launcher(nx, nu)) return false;  // (widths are padded to multiples of 16)
  const int env = env_int("TAMPC_TC", -1);
  if (env == 0) return false;
  if (env == 1 || !has_tuned_variant) return true;
  // measured on B200 (profiles/r2_tc_sweep.jsonl): at widths >= 128 the tcgen05 kernel wins at every K (H=256: 0.53 vs 2.03 ms
  // at K=1024, 29.6 vs 113 ms at K=1M; H=128: 0.22 vs 0.23 ms at K=1024, 12.3 vs 23.7 ms at 1M); at 32 / 64 the register-
  // chained mma.sync kernels win (tcgen05.mma issue costs ~150 cycles whatever its N)
  return (h1 >= 128 || h2 >= 128) && h->cfg.num_local >= env_int("TAMPC_TC_MIN_K", 1);
}

// placeholder variants of the runtime-shaped kernel (generic_rollout.cuh)
const ModelVariant* variant_generic(int kind) {
  static const ModelVariant mlp = [] { ModelVariant m{}; m.name = "mlp_generic"; m.kind = TAMPC_DYN_MLP; return m; }();
  static const ModelVariant rex = [] { ModelVariant m{}; m.name = "rex_generic"; m.kind = TAMPC_DYN_REX; return m; }();
  return kind == TAMPC_DYN_REX ? &rex : &mlp;
}

template <typename NT>
void pack_gen_net(std::vector<NT>& blob, const std::vector<HostLayer>& m, GenNet& g) {
  g.n = (int)m.size();
  g.dims[0] = m[0].in;
  for (size_t l = 0; l < m.size(); ++l) {
    g.off[l] = (int)blob.size();
    g.dims[l + 1] = m[l].out;
    pack_layer<NT>(blob, m[l]);
  }
}

bool gen_widths_ok(const std::vector<HostLayer>& m) {
  for (const HostLayer& L : m)
    if (L.in > kGenMaxWidth || L.out > kGenMaxWidth) return false;
  return true;
}

// TAMPC_GENERIC=1: the runtime-shaped kernel even where a tuned one exists (tests compare the two bit for bit)
bool generic_forced() { return env_int("TAMPC_GENERIC", 0) != 0; }

// does the tuned variant have a kernel for the handle's precision?  (if not, the runtime-shaped kernel takes the model;
// a GP residual needs the variant's own (sample, replica) kernels, so there the variant stays and install_model decides)
bool variant_serves(const tampc_mppi* h, const ModelVariant* v) {
  if (h->gp_on) return true;
  switch (h->cfg.precision) {
    case TAMPC_PREC_F64: return v->bytes_mma_f64 != 0 || v->f64_s1 != nullptr;
    case TAMPC_PREC_TF32X3: return v->bytes_mma_tf32 != 0;
    case TAMPC_PREC_F16X3: return v->bytes_mma_f16 != 0;
    default: return v->f64_s1 != nullptr;
  }
}

int pack_model(tampc_mppi* h, const tampc_model_desc& d, const double* p, size_t n, PackedModel& out) {
  const int nx = d.nx, nu = d.nu;
  if (nx != h->cfg.nx || nu != h->cfg.nu) { h->err = "model nx/nu differ from the handle's"; return TAMPC_ERR_INVALID; }
  if (!(d.leaky_slope >= 0.0 && d.leaky_slope <= 1.0)) { h->err = "leaky_slope must be in [0,1]"; return TAMPC_ERR_UNSUPPORTED; }
  const ModelVariant* all[] = {variant_rex_push(), variant_rex_peg(), variant_mlp_5_3_32(), variant_mlp_5_3_64(), variant_mlp_5_3_128(),
                               variant_pendulum()};
  char buf[256];
  if (d.kind == TAMPC_DYN_PENDULUM) {
    if (nx != 2 || nu != 1) { h->err = "pendulum dynamics need nx=2, nu=1"; return TAMPC_ERR_INVALID; }
    out.variant = variant_pendulum();
    out.f32.assign(8, 0.f);
    out.f64.assign(8, 0.0);
    for (int i = 0; i < 6; ++i) { out.f32[i] = (float)d.pendulum[i]; out.f64[i] = d.pendulum[i]; }
    return TAMPC_OK;
  }
  if (d.kind == TAMPC_DYN_MLP) {
    std::vector<HostLayer> net;
    int rc = read_mlp(h, d.net, p, n, net);
    if (rc) return rc;
    if (net[0].in != nx + nu || net.back().out != nx) {
      h->err = "MLP dynamics must map (nx+nu) -> ... -> nx";
      return TAMPC_ERR_UNSUPPORTED;
    }
    std::vector<double> s, o;
    if ((rc = read_affine(h, d.aff_in, p, n, nx + nu, s, o))) return rc;
    fold_input_affine(net[0], s, o);
    for (int i = 0; i < nx + nu; ++i) { h->net_in_scale[i] = s[i]; h->net_in_off[i] = o[i]; }
    if ((rc = read_affine(h, d.aff_out, p, n, nx, s, o))) return rc;
    for (int i = 0; i < nx; ++i) h->net_out_scale[i] = s[i];
    {  // dx = (y - o)/s
      std::vector<double> si(nx), oi(nx);
      for (int i = 0; i < nx; ++i) { si[i] = 1.0 / s[i]; oi[i] = -o[i] / s[i]; }
      fold_output_affine(net.back(), si, oi);
    }
    for (const ModelVariant* v : all) {
      if (net.size() == 3 && v->kind == TAMPC_DYN_MLP && v->nx == nx && v->nu == nu && net[0].out == v->dims[0] && net[1].out == v->dims[1]) out.variant = v;
    }
    if (out.variant && (generic_forced() || !variant_serves(h, out.variant))) out.variant = nullptr;
    if (!generic_forced() && tc_mlp_wanted(h, net, nx, nu, out.variant != nullptr)) {
      pack_tc_mlp(net, nx, nu, out);
      out.variant = variant_tc_generic();
      return TAMPC_OK;
    }
    if (!out.variant) {
      // any other depth / widths / precision: the runtime-shaped kernel
      if (!gen_widths_ok(net)) {
        snprintf(buf, sizeof buf, "MLP dynamics with a layer wider than %d", kGenMaxWidth);
        h->err = buf;
        return TAMPC_ERR_UNSUPPORTED;
      }
      out.generic = true;
      out.gen_f32.kind = out.gen_f64.kind = TAMPC_DYN_MLP;
      pack_gen_net<float>(out.f32, net, out.gen_f32.net);
      pack_gen_net<double>(out.f64, net, out.gen_f64.net);
      out.variant = variant_generic(TAMPC_DYN_MLP);
      return TAMPC_OK;
    }
    if (out.variant->f64_s1) {  // shapes with FMA-path kernels
      pack_mlp<float>(out.f32, net);
      pack_mlp<double>(out.f64, net);
    }
    if (out.variant->bytes_mma_f64) pack_mlp_mma(out.mma_f64, net, 1, true);
    if (out.variant->bytes_mma_tf32) pack_mlp_mma(out.mma_tf32, net, 1, false);
    if (out.variant->bytes_mma_f16) pack_mlp_mma(out.mma_f16, net, 1, 2);
    return TAMPC_OK;
  }
  if (d.kind != TAMPC_DYN_REX) { h->err = "unknown dynamics kind"; return TAMPC_ERR_INVALID; }
  std::vector<HostLayer> enc, dyn, dec;
  int rc;
  if ((rc = read_mlp(h, d.encoder, p, n, enc)) || (rc = read_mlp(h, d.net, p, n, dyn)) || (rc = read_mlp(h, d.decoder, p, n, dec))) return rc;
  const int nv = d.nv, ne = d.ne;
  if (enc[0].in != nx + nu || dyn[0].in != enc.back().out || dyn.back().out != nv || dec[0].in != ne || dec.back().out != nx * nv) {
    h->err = "REX dynamics must be three MLPs: (nx+nu)->..->nz, nz->..->nv, ne->..->nx*nv";
    return TAMPC_ERR_UNSUPPORTED;
  }
  const int nz = enc.back().out;
  std::vector<double> s, o;
  // z' = z*s + o  after the encoder
  if ((rc = read_affine(h, d.aff_in, p, n, nz, s, o))) return rc;
  fold_output_affine(enc.back(), s, o);
  // v = (v' - o)/s after the latent dynamics
  if ((rc = read_affine(h, d.aff_out, p, n, nv, s, o))) return rc;
  if (nv > kMaxNx) { h->err = "latent output wider than TAMPC_MAX_NX"; return TAMPC_ERR_UNSUPPORTED; }
  for (int i = 0; i < nv; ++i) h->net_out_scale[i] = s[i];
  {
    std::vector<double> si(nv), oi(nv);
    for (int i = 0; i < nv; ++i) { si[i] = 1.0 / s[i]; oi[i] = -o[i] / s[i]; }
    fold_output_affine(dyn.back(), si, oi);
  }
  // e = Wx x + bx feeds the decoder's first Linear with no activation in between: compose them
  if (d.ext_w_off < 0 || d.ext_b_off < 0 || (size_t)d.ext_w_off + (size_t)ne * nx > n || (size_t)d.ext_b_off + ne > n) {
    h->err = "extractor offsets outside the params array";
    return TAMPC_ERR_INVALID;
  }
  {
    const double* Wx = p + d.ext_w_off;
    const double* bx = p + d.ext_b_off;
    HostLayer L;
    L.in = nx;
    L.out = dec[0].out;
    L.W.assign((size_t)L.out * nx, 0.0);
    L.b = dec[0].b;
    for (int r = 0; r < L.out; ++r) {
      for (int e = 0; e < ne; ++e) {
        const double w = dec[0].W[(size_t)r * ne + e];
        L.b[r] += w * bx[e];
        for (int c = 0; c < nx; ++c) L.W[(size_t)r * nx + c] += w * Wx[(size_t)e * nx + c];
      }
    }
    dec[0] = L;
  }
  // dx_i = (sum_j C[i][j] v_j - o_i)/s_i : scale decoder output rows i*nv+j by 1/s_i, keep d0_i = -o_i/s_i
  if ((rc = read_affine(h, d.aff_y, p, n, nx, s, o))) return rc;
  std::vector<double> d0(nx);
  {
    std::vector<double> si((size_t)nx * nv), oi((size_t)nx * nv, 0.0);
    for (int i = 0; i < nx; ++i) {
      d0[i] = -o[i] / s[i];
      for (int j = 0; j < nv; ++j) si[(size_t)i * nv + j] = 1.0 / s[i];
    }
    fold_output_affine(dec.back(), si, oi);
  }
  for (const ModelVariant* v : all) {
    if (enc.size() == 3 && dyn.size() == 3 && dec.size() == 3 && v->kind == TAMPC_DYN_REX && v->nx == nx && v->nu == nu && enc[0].out == v->dims[0] &&
        enc[1].out == v->dims[1] && nz == v->dims[2] && dyn[0].out == v->dims[3] && dyn[1].out == v->dims[4] && nv == v->dims[5] &&
        dec[0].out == v->dims[6] && dec[1].out == v->dims[7])
      out.variant = v;
  }
  if (out.variant && (generic_forced() || !variant_serves(h, out.variant))) out.variant = nullptr;
  if (!out.variant) {
    // other depths / widths: the runtime-shaped kernel
    if (!gen_widths_ok(enc) || !gen_widths_ok(dyn) || !gen_widths_ok(dec)) {
      snprintf(buf, sizeof buf, "REX dynamics with a layer wider than %d", kGenMaxWidth);
      h->err = buf;
      return TAMPC_ERR_UNSUPPORTED;
    }
    out.generic = true;
    for (int dbl = 0; dbl < 2; ++dbl) {
      GenDims& g = dbl ? out.gen_f64 : out.gen_f32;
      g.kind = TAMPC_DYN_REX;
      g.nv = nv;
      if (dbl) {
        pack_gen_net<double>(out.f64, enc, g.enc);
        pack_gen_net<double>(out.f64, dyn, g.net);
        pack_gen_net<double>(out.f64, dec, g.dec);
        g.off_d0 = (int)out.f64.size();
      } else {
        pack_gen_net<float>(out.f32, enc, g.enc);
        pack_gen_net<float>(out.f32, dyn, g.net);
        pack_gen_net<float>(out.f32, dec, g.dec);
        g.off_d0 = (int)out.f32.size();
      }
    }
    const int d0g = (nx + 3) & ~3;
    for (int i = 0; i < d0g; ++i) {
      out.f32.push_back(i < nx ? (float)d0[i] : 0.f);
      out.f64.push_back(i < nx ? d0[i] : 0.0);
    }
    out.variant = variant_generic(TAMPC_DYN_REX);
    return TAMPC_OK;
  }
  pack_mlp<float>(out.f32, enc);
  pack_mlp<float>(out.f32, dyn);
  pack_mlp<float>(out.f32, dec);
  pack_mlp<double>(out.f64, enc);
  pack_mlp<double>(out.f64, dyn);
  pack_mlp<double>(out.f64, dec);
  const int d0p = (nx + 3) & ~3;
  for (int i = 0; i < d0p; ++i) {
    out.f32.push_back(i < nx ? (float)d0[i] : 0.f);
    out.f64.push_back(i < nx ? d0[i] : 0.0);
  }
  if (out.variant->bytes_mma_f64) {
    for (int kind = 0; kind < 3; ++kind) {  // 0 = 3xTF32, 1 = float64, 2 = 3xFP16
      if (kind == 2 && !out.variant->bytes_mma_f16) continue;
      std::vector<unsigned char>& blob = kind == 1 ? out.mma_f64 : (kind == 2 ? out.mma_f16 : out.mma_tf32);
      pack_mlp_mma(blob, enc, 1, kind);
      pack_mlp_mma(blob, dyn, 1, kind);
      pack_mlp_mma(blob, dec, (nx * nv + 7) / 8, kind);
      for (int i = 0; i < 8; ++i) {
        const double v = i < nx ? d0[i] : 0.0;
        if (kind == 1) append<double>(blob, v); else append<float>(blob, (float)v);
      }
    }
  }
  return TAMPC_OK;
}

// ------------------------------------------------------------------------------------------------

// Launch shape.  FMA kernels: S samples per thread.  mma kernels: NM m-tiles (8 or 16 samples each) per warp.
// Small K is latency-bound: spread samples over as many warps / SMs as possible; large K is throughput-bound:
// more samples per warp re-use each weight fragment.  TAMPC_SAMPLES_PER_THREAD / TAMPC_MMA_NM / TAMPC_BLOCK override.
struct LaunchPlan {
  rollout_launch_fn fn = nullptr;
  const char* name = "none";  // kernel family + arithmetic that will run (tampc_mppi_kernel_name)
  int block = 128;
  // > 0: the kernel can run the fused softmax update (fused_reduce.cuh); number of partial slots (warps) it uses.
  // Opt-in (TAMPC_FUSE=1): measured on B200 it saves ~3 us per command when the softmax is peaked (lambda = 0.01
  // with pushing-scale costs) but its single merging warp is several times slower than the separate
  // partials/combine kernels when most weights are non-zero (peg-scale costs), so the default is the robust path.
  int fuse_warps = 0;
};

constexpr int kMaxFusedWarps = 160 * 4;  // capacity of the fused-update scratch: one warp per SM sub-partition (latency-bound sizes)

LaunchPlan choose_launch(const tampc_mppi* h, int K) {
  const ModelVariant* v = h->variant;
  LaunchPlan p;
  const int SM = h->n_sm;  // cudaDeviceProp::multiProcessorCount (148 on B200); the thresholds below are per SM
  const int eb = env_int("TAMPC_BLOCK", 0);
  if (h->use_generic) {
    p.fn = h->gen_fn;
    p.name = h->gen_name;
    p.block = K >= 2 * SM * 128 ? 128 : (K >= 2 * SM * 64 ? 64 : 32);
    if (eb == 32 || eb == 64 || eb == 128) p.block = eb;
    return p;
  }
  if (h->use_tc) {
    p.fn = h->tc_fn;
    p.block = tc::kThreads;
    p.name = "tc_mlp_rollout_kernel 3xTF32 (tcgen05.mma kind::tf32, TMEM operands, TMA weight ring)";
    return p;
  }
  if (h->gp_on) {
    // one thread per (sample, replica), FMA-layout networks; float32 networks for every precision but F64
    if (h->img_f64) { p.fn = v->gp.f64; p.name = "rollout_gp_kernel f64 (DFMA)"; }
    else { p.fn = v->gp.mixed; p.name = "rollout_gp_kernel f32 networks (FFMA2), f64 state + cost + GP posterior"; }
    p.block = 128;
    return p;
  }
  if (h->use_mma) {
    const bool f64 = h->cfg.precision == TAMPC_PREC_F64, f16 = h->cfg.precision == TAMPC_PREC_F16X3;
    const rollout_launch_fn* fmma = f64 ? v->f64_mma : (f16 ? v->f16_mma : v->tf32_mma);
    const rollout_launch_fn fwide = f16 ? v->f16_mma_wide : v->tf32_mma_wide, fsplit = f16 ? v->f16_split : v->tf32_split;
    const rollout_launch_fn fpipe = f16 ? nullptr : v->tf32_pipe;
    const int mt = f64 ? 8 : 16, max_idx = f64 ? 2 : 1;
    // Resident warps per SM are bounded by shared memory (image ~40 KB per CTA + ~14 KB per warp): about 4 one-warp
    // CTAs, 3 two-warp CTAs or 2 four-warp CTAs.  While everything fits in one wave the kernel is latency-bound and
    // the smallest tile (most warps) wins; a second wave costs a full kernel time, so beyond one wave of the
    // smallest tile move to the next tile size (measured on B200, tools/perf_sweep.py).
    const bool big = h->img_bytes > 100 * 1024;  // wide networks: one CTA per SM, always four warps sharing the image
    const int cap32 = big ? 0 : SM * 4, cap64 = big ? 0 : SM * 3 * 2, cap128 = big ? SM * 4 : SM * 2 * 4;
    int idx = 0;  // NM = 1 << idx
    while (idx < max_idx && fmma[idx + 1] != nullptr && (K + (mt << idx) - 1) / (mt << idx) > cap128) ++idx;
    const int en = env_int("TAMPC_MMA_NM", 0);
    if (en == 1) idx = 0; else if (en == 2) idx = 1; else if (en == 4 && f64) idx = 2;
    const int warps = (K + (mt << idx) - 1) / (mt << idx);
    p.block = warps <= cap32 ? 32 : (warps <= cap64 ? 64 : 128);
    p.fn = fmma[idx];
    p.name = f64 ? "rollout_mma_kernel f64 (DMMA m8n8k4)" : (f16 ? "rollout_mma_kernel 3xFP16 (HMMA m16n8k16)" : "rollout_mma_kernel 3xTF32 (HMMA m16n8k8)");
    if (eb == 32 || eb == 64 || eb == 128) p.block = eb;
    // throughput-bound 3xTF32: eight warps per CTA and two CTAs per SM give 16 resident warps instead of 12.
    // Measured on B200 (push, T=40): a full wave takes 1.18x as long but holds 1.33x the samples (1M: 5.98 -> 5.68 ms),
    // so it is used when it does not add a wave (K=131072: 0.83 vs 0.78 ms).  TAMPC_WIDE=0/1 forces the choice.
    // one balanced wave: when 12 warps per SM are not enough for a single wave but 16 are (56 832 < K <= 75 776), one CTA
    // per SM with ceil(warps / SM) warps gives every SM the same work (K=65536: 14 warps each instead of 16 on 108 SMs
    // and 8 on 40).  TAMPC_SM_WAVE=0 disables.
    // ... and seven tiles per sub-partition (three two-tile warps + one one-tile warp) while 448 samples per SM are
    // enough: the wave then lasts as long as 3.5 instead of 4 two-tile warps per sub-partition.  TAMPC_MIX=0 disables.
    const rollout_launch_fn fmix = f16 ? v->f16_mma_mix : v->tf32_mma_mix;
    if (!f64 && idx == 1 && !big && fmix && eb == 0 && warps > SM * 12 && K <= SM * 448 && env_int("TAMPC_MIX", 1) && env_int("TAMPC_SM_WAVE", 1)) {
      p.fn = fmix;
      p.name = f16 ? "rollout_mma_mix_kernel 3xFP16 (HMMA m16n8k16)" : "rollout_mma_mix_kernel 3xTF32 (HMMA m16n8k8)";
      p.block = 512;
      return p;
    }
    const rollout_launch_fn fsm = f16 ? v->f16_mma_sm : v->tf32_mma_sm;
    if (!f64 && idx == 1 && !big && fsm && eb == 0 && warps > SM * 12 && warps <= SM * 16 && env_int("TAMPC_SM_WAVE", 1)) {
      p.fn = fsm;
      p.block = 32 * ((warps + SM - 1) / SM);
      return p;
    }
    // more than one wave of sixteen warps per SM: one persistent 16-warp CTA per SM that walks the 32-sample units
    // slot-major, so that a partial last pass spreads over all sub-partitions.  Measured on B200 against the grid launches
    // below (profiles/r2o_persist_*.jsonl, rollout kernel, T=40): push K=131 072 0.782 -> 0.723 ms, 262 144 1.527 -> 1.444,
    // 1M 5.57 -> 5.59; synthetic H=32 K=131 072 0.462 -> 0.348 ms, 262 144 0.799 -> 0.690, 1M 2.63 -> 2.60.  TAMPC_PERSIST=0/1.
    const rollout_launch_fn fpersist = f16 ? v->f16_mma_persist : v->tf32_mma_persist;
    if (!f64 && idx == 1 && !big && fpersist && eb == 0 && warps > SM * 16) {
      const int pe = env_int("TAMPC_PERSIST", -1);
      if (pe != 0) {
        p.fn = fpersist;
        p.name = f16 ? "rollout_mma_persist_kernel 3xFP16 (HMMA m16n8k16)" : "rollout_mma_persist_kernel 3xTF32 (HMMA m16n8k8)";
        p.block = 512;
        return p;
      }
    }
    if (!f64 && idx == 1 && !big && fwide && eb == 0) {
      const int wide_env = env_int("TAMPC_WIDE", -1);
      const int ctas_d = (warps + 3) / 4, ctas_w = (warps + 7) / 8;
      const bool wide = wide_env >= 0 ? wide_env != 0 : (ctas_d > SM * 3 && ctas_w <= SM * 2) || ctas_w >= SM * 2 * 4;
      if (wide) {
        p.fn = fwide;
        p.block = 256;
      }
    }
    {
      const int wpb = p.block / 32, grid = (warps + wpb - 1) / wpb;
      if (grid * wpb <= kMaxFusedWarps && env_int("TAMPC_FUSE", 0)) p.fuse_warps = grid * wpb;
    }
    // latency-bound sizes: the two-warp pipelined kernel (dynamics warp + cost warp per 16-sample tile)
    // (measured on B200: wins while at most two 16-sample tiles share an SM, i.e. K <= 4736; above that the
    //  single-warp kernel's lower footprint wins.  TAMPC_PIPE=0/1 forces the choice.)
    // latency-bound sizes of the RexExtract shapes: two warps per 16-sample tile split by role (mma_rollout.cuh).
    // Measured on B200 (push, T=40; rollout kernel, split vs one warp per tile): K=512 64 vs 74 us, 4096 67 vs 78,
    // 8192 86 vs 112, 12288 117 vs 141; at 16384 (more than five tiles per SM) 161 vs 145, so the single-warp
    // kernel takes over.  TAMPC_SPLIT=0/1 forces the choice.
    {
      const int split_env = env_int("TAMPC_SPLIT", -1);
      const bool split = !big && (split_env >= 0 ? split_env != 0 : (K + 15) / 16 <= 5 * SM + 28);
      if (!f64 && idx == 0 && fsplit && split && en == 0) {
        // a third warp takes the cost off the decoder warp.  Measured on B200 (push, T=40, 3xTF32): K=500 64 -> 60 us,
        // 4736 67.5 -> 64, 8192 86.9 -> 77.2, 12288 117 -> 117 (peg K=8192, T=10: 27.9 -> 23.8).  TAMPC_SPLIT3=0/1 forces.
        const rollout_launch_fn fsplit3 = f16 ? v->f16_split3 : v->tf32_split3;
        const int s3_env = env_int("TAMPC_SPLIT3", -1);
        const bool three = fsplit3 && (s3_env >= 0 ? s3_env != 0 : true);
        p.fn = three ? fsplit3 : fsplit;
        p.name = f16 ? "rollout_mma_split_kernel 3xFP16 (HMMA m16n8k16)" : "rollout_mma_split_kernel 3xTF32 (HMMA m16n8k8)";
        p.block = three ? 192 : 128;
        p.fuse_warps = 0;
        return p;
      }
    }
    const int tiles = (K + 15) / 16, pipe_env = env_int("TAMPC_PIPE", -1);
    const bool pipe = !big && (pipe_env >= 0 ? pipe_env != 0 : tiles <= 2 * SM);
    if (!f64 && idx == 0 && fpipe && pipe) {
      p.fn = fpipe;
      p.name = "rollout_mma_pipe_kernel 3xTF32 (HMMA m16n8k8)";
      p.block = 64;
      p.fuse_warps = env_int("TAMPC_FUSE", 0) ? tiles : 0;
      return p;
    }
  } else {
    const bool f64 = h->cfg.precision == TAMPC_PREC_F64;
    const bool has_s2 = !f64 && v->mixed_s2 != nullptr;
    int S = (has_s2 && K >= 2 * SM * 128 * 2) ? 2 : 1;
    const int es = env_int("TAMPC_SAMPLES_PER_THREAD", 0);
    if (es == 1 || (es == 2 && has_s2)) S = es;
    const int threads = (K + S - 1) / S;
    p.block = threads >= 2 * SM * 128 ? 128 : (threads >= 2 * SM * 64 ? 64 : 32);
    switch (h->cfg.precision) {
      case TAMPC_PREC_F64: p.fn = v->f64_s1; p.name = "rollout_kernel f64 (DFMA)"; break;
      case TAMPC_PREC_MIXED:
      case TAMPC_PREC_TF32X3:
      case TAMPC_PREC_F16X3:
        // shapes without network layers (analytic pendulum) have nothing for the tensor cores: the FMA kernel runs
        p.fn = S == 2 ? v->mixed_s2 : v->mixed_s1;
        p.name = "rollout_kernel f32 arithmetic (FFMA2), f64 state+cost";
        break;
      default: p.fn = S == 2 ? v->f32_s2 : v->f32_s1; p.name = "rollout_kernel f32 (FFMA2)"; break;
    }
  }
  if (eb == 32 || eb == 64 || eb == 128) p.block = eb;
  return p;
}

// (re)build the CostS / SamplerS parts of the shared-memory image in the types the selected kernels use
int prep_image(tampc_mppi* h) {
  if (!h->has_model || !h->has_cost) return TAMPC_OK;
  unsigned char* img = reinterpret_cast<unsigned char*>(h->d_weights);
  const int prec = h->cfg.precision;
  const int sd = h->use_mma ? h->img_sampler_d_off : -1;
  if (prec == TAMPC_PREC_F64 || h->gp_on) prep_image_kernel<double, double><<<1, 128, 0, h->stream>>>(img, h->img_cost_off, h->img_sampler_off, sd, h->d_cost, h->d_sampler);
  else if (prec == TAMPC_PREC_MIXED || ((prec == TAMPC_PREC_TF32X3 || prec == TAMPC_PREC_F16X3) && !h->use_mma))
    prep_image_kernel<double, float><<<1, 128, 0, h->stream>>>(img, h->img_cost_off, h->img_sampler_off, sd, h->d_cost, h->d_sampler);
  else prep_image_kernel<float, float><<<1, 128, 0, h->stream>>>(img, h->img_cost_off, h->img_sampler_off, sd, h->d_cost, h->d_sampler);
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return TAMPC_OK;
}

SampleSrc make_src(const tampc_mppi* h, int mode, uint64_t call, int T) {
  SampleSrc s{};
  s.mode = mode;
  s.noise = h->d_noise;
  s.actions = h->d_noise;
  s.seed = h->cfg.seed;
  s.call = call;
  s.sample_offset = h->cfg.sample_offset;
  s.k_global = h->cfg.num_samples;
  s.T = T;
  return s;
}

int check_ready(tampc_mppi* h) {
  if (!h->has_model) { h->err = "set_model has not been called"; return TAMPC_ERR_STATE; }
  if (!h->has_cost) { h->err = "set_cost has not been called"; return TAMPC_ERR_STATE; }
  if (!h->has_nominal) { h->err = "set_nominal has not been called"; return TAMPC_ERR_STATE; }
  return TAMPC_OK;
}

// online GP residual: kernel arguments of the (sample, replica) kernels; consumes injected draws if armed (one-shot)
int apply_gp_args(tampc_mppi* h, RolloutArgs& a, int T) {
  if (!h->gp_on) return TAMPC_OK;
  a.gp = h->d_gp;
  a.gp_bytes = h->gp_bytes;
  a.M = h->gp.rollout_samples;
  a.var_cost = h->gp.rollout_var_cost;
  a.var_discount = h->gp.rollout_var_discount;
  // get_rollouts / predict / rollout_costs of given actions are sampled too in the reference (they go through the same dynamics): a Philox stream
  // of their own, advanced per call
  if (a.src.mode == kModeNominal || a.src.mode == kModeActions) a.src.call = (1ull << 62) + h->gp_aux_calls++;
  if (h->gp_eps_armed) {
    if (T != h->gp_eps_T) { h->err = "the injected GP draws were sized for another horizon"; return TAMPC_ERR_INVALID; }
    a.gp_eps = h->d_gp_eps;
    h->gp_eps_armed = false;
  }
  return TAMPC_OK;
}

// fuse: 0 = rollout only; 1 / 2 = ask for the fused softmax update (finalise / rank partial).  Returns in *fused
// whether the selected kernel did it (small-K tensor-core kernels) or the separate reduction kernels must follow.
int launch_rollout_for(tampc_mppi* h, int mode, int shift, const double* d_U, int T, uint64_t call, bool keep_states, int fuse = 0,
                       tampc_partial* partial_out = nullptr, bool* fused = nullptr) {
  RolloutArgs a{};
  a.image = h->d_weights;
  a.image_bytes = h->img_bytes;
  a.tc = h->tc_dims;
  a.gen = h->gen;
  a.tc_w2 = h->d_tc_w2;
  a.tc_prof = h->d_tc_prof;
  a.slope = h->model.leaky_slope;
  a.cost = h->d_cost;
  a.sampler = h->d_sampler;
  a.U = d_U;
  a.shift = shift;
  memcpy(a.state, h->state_host, sizeof a.state);
  a.src = make_src(h, mode, call, T);
  a.K_local = h->cfg.num_local;
  a.cost_out = h->d_costs;
  a.states_out = keep_states ? h->d_states : nullptr;
  a.wrap_mask = h->model.wrap_mask;
  a.has_guard = h->model.has_bad_state_guard;
  a.bad_norm = h->model.bad_state_norm;
  {
    int rc = apply_gp_args(h, a, T);
    if (rc) return rc;
  }
  const LaunchPlan plan = choose_launch(h, a.K_local);
  if (!plan.fn) { h->err = "no kernel compiled for this precision"; return TAMPC_ERR_UNSUPPORTED; }
  if (fused) *fused = false;
  if (fuse && plan.fuse_warps > 0 && plan.fuse_warps <= kMaxFusedWarps) {
    a.fuse = fuse;
    a.phdr = h->d_phdr;
    a.pwsum = h->d_pwsum;
    a.ticket = h->d_ticket;
    a.U_out = h->d_U[h->cur ^ 1];
    a.action_out = h->d_action;
    a.stats_out = h->d_stats;
    a.partial_out = partial_out ? partial_out : h->d_rank_partial;
    if (fuse == 1) {
      a.host_ll = h->d_map;
      a.seq = ++h->seq;
    }
    if (fused) *fused = true;
  }
  CUDA_TRY(h, plan.fn(a, plan.block, h->stream));
  h->launches++;
  return TAMPC_OK;
}

int ensure_states(tampc_mppi* h) {
  if (!h->d_states)
    CUDA_TRY(h, cudaMalloc(&h->d_states, sizeof(double) * (size_t)h->cfg.num_local * h->cfg.max_horizon * h->cfg.nx));
  return TAMPC_OK;
}

int ensure_noise(tampc_mppi* h) {
  if (!h->d_noise)
    CUDA_TRY(h, cudaMalloc(&h->d_noise, sizeof(double) * (size_t)h->cfg.num_local * h->cfg.max_horizon * h->cfg.nu));
  return TAMPC_OK;
}

CombineArgs base_combine_args(tampc_mppi* h) {
  CombineArgs c{};
  c.sampler = h->d_sampler;
  c.U = h->d_U[h->cur];
  c.T = h->T;
  c.nu = h->cfg.nu;
  c.U_out = h->d_U[h->cur ^ 1];
  c.action_out = h->d_action;
  c.stats_out = h->d_stats;
  return c;
}

ExchangeArgs make_exchange_args(tampc_mppi* h);

// The tail launches are programmatic dependents of the rollout launch in front of them (TAMPC_PDL=0: plain stream
// order): their CTAs are placed while the rollout still runs and wait in grid_dep_wait() for its completion, so the
// launch latency and the staging of the sampler constants are off the command's critical path.
template <typename... KArgs, typename... Args>
cudaError_t launch_dependent(void (*kern)(KArgs...), int grid, int block, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const int pdl = env_int("TAMPC_PDL", 1);
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int NU>
cudaError_t launch_partials_exchange(const ReduceArgs& r, const CombineArgs& c, const ExchangeArgs& e, unsigned int* ticket, int n_chunks, cudaStream_t s) {
  return launch_dependent(partials_exchange_kernel<NU>, n_chunks, kCombineThreads, s, r, c, e, ticket);
}

template <int NU>
cudaError_t launch_partials_combine(const ReduceArgs& r, const CombineArgs& c, unsigned int* ticket, int n_chunks, cudaStream_t s) {
  return launch_dependent(partials_combine_kernel<NU>, n_chunks, kCombineThreads, s, r, c, ticket);
}

template <int NU>
void launch_partials(const ReduceArgs& r, int n_chunks, cudaStream_t s) {
  partials_kernel<NU><<<n_chunks, kReduceThreads, 0, s>>>(r);
}

// rollout of the shard + its reduction.  fuse = 1 (finalise) / 2 (rank partial to partial_out).  *done tells the
// caller whether the softmax update already happened inside the rollout launch (small-K tensor-core kernels) or the
// CTA partials are waiting in d_partials for run_combine.
int run_rollout_and_partials(tampc_mppi* h, int mode, bool keep_states, int fuse, tampc_partial* partial_out, bool* done) {
  // fuse = 3: like 2 (rank partial), and the peer exchange + world merge join the same launch when the fused tail applies
  // (h->tail_exchanged tells the caller whether they did)
  const bool want_exchange = fuse == 3;
  if (want_exchange) fuse = 2;
  h->tail_exchanged = false;
  const int T = h->T;
  const double* U = h->d_U[h->cur];
  h->call++;
  bool fused = false;
  int rc = launch_rollout_for(h, mode, /*shift=*/1, U, T, h->call, keep_states, fuse, partial_out, &fused);
  if (rc) return rc;
  h->last_mode = mode;
  h->last_call = h->call;
  h->last_T = T;
  h->last_U = h->cur;
  h->last_valid = true;
  h->states_valid = keep_states;
  *done = fused;
  if (fused) {
    if (fuse == 1) h->cur ^= 1;
    return TAMPC_OK;
  }
  ReduceArgs r{};
  r.cost = h->d_costs;
  r.sampler = h->d_sampler;
  r.U = U;
  r.src = make_src(h, mode, h->call, T);
  r.K_local = h->cfg.num_local;
  r.chunk = h->chunk;
  r.nu = h->cfg.nu;
  r.partials = h->d_partials;
  // latency path: both reduction stages in one launch while a single merging CTA suffices (TAMPC_FUSE_TAIL=0: separate)
  if (want_exchange && h->n_chunks <= kCombineThreads && env_int("TAMPC_FUSE_TAIL", 1)) {
    CombineArgs c = base_combine_args(h);
    c.partials = h->d_partials;
    c.n_partials = h->n_chunks;
    c.group = h->n_chunks;
    c.solo = 1;
    c.finalize = 0;
    c.partial_out = h->d_rank_partial;
    const ExchangeArgs e = make_exchange_args(h);
    cudaError_t le;
    switch (h->cfg.nu) {
      case 1: le = launch_partials_exchange<1>(r, c, e, h->d_ticket, h->n_chunks, h->stream); break;
      case 2: le = launch_partials_exchange<2>(r, c, e, h->d_ticket, h->n_chunks, h->stream); break;
      case 3: le = launch_partials_exchange<3>(r, c, e, h->d_ticket, h->n_chunks, h->stream); break;
      default: le = launch_partials_exchange<4>(r, c, e, h->d_ticket, h->n_chunks, h->stream); break;
    }
    CUDA_TRY(h, le);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
    h->cur ^= 1;
    h->tail_exchanged = true;
    *done = true;
    return TAMPC_OK;
  }
  if (fuse && h->n_chunks <= kCombineThreads && env_int("TAMPC_FUSE_TAIL", 1)) {
    CombineArgs c = base_combine_args(h);
    c.partials = h->d_partials;
    c.n_partials = h->n_chunks;
    c.group = h->n_chunks;
    c.solo = 1;
    c.finalize = fuse == 1;
    c.partial_out = partial_out ? partial_out : h->d_rank_partial;
    if (fuse == 1) {
      c.host_ll = h->d_map;
      c.seq = ++h->seq;
    }
    cudaError_t le;
    switch (h->cfg.nu) {
      case 1: le = launch_partials_combine<1>(r, c, h->d_ticket, h->n_chunks, h->stream); break;
      case 2: le = launch_partials_combine<2>(r, c, h->d_ticket, h->n_chunks, h->stream); break;
      case 3: le = launch_partials_combine<3>(r, c, h->d_ticket, h->n_chunks, h->stream); break;
      default: le = launch_partials_combine<4>(r, c, h->d_ticket, h->n_chunks, h->stream); break;
    }
    CUDA_TRY(h, le);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
    if (fuse == 1) h->cur ^= 1;
    *done = true;
    return TAMPC_OK;
  }
  switch (h->cfg.nu) {
    case 1: launch_partials<1>(r, h->n_chunks, h->stream); break;
    case 2: launch_partials<2>(r, h->n_chunks, h->stream); break;
    case 3: launch_partials<3>(r, h->n_chunks, h->stream); break;
    default: launch_partials<4>(r, h->n_chunks, h->stream); break;
  }
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return TAMPC_OK;
}

// Merge n partials.  More than one group of kCombineThreads is first reduced group-wise into d_partials2 (a
// deterministic two-level tree), then a single CTA finishes: finalize=1 writes the new U / action / stats,
// finalize=0 writes one merged partial (this rank's contribution) to partial_out.
int run_combine(tampc_mppi* h, const tampc_partial* partials, int n, int finalize, tampc_partial* partial_out = nullptr) {
  CombineArgs c = base_combine_args(h);
  while (n > kCombineThreads) {
    const int groups = (n + kCombineThreads - 1) / kCombineThreads;
    if (groups > h->n_partials2) { h->err = "internal: too many partials for the combine scratch"; return TAMPC_ERR_INVALID; }
    c.partials = partials;
    c.n_partials = n;
    c.group = kCombineThreads;
    c.finalize = 0;
    c.partial_out = h->d_partials2;
    combine_kernel<<<groups, kCombineThreads, 0, h->stream>>>(c);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
    partials = h->d_partials2;
    n = groups;
  }
  c.partials = partials;
  c.n_partials = n;
  c.group = n;
  c.finalize = finalize;
  c.partial_out = partial_out ? partial_out : h->d_rank_partial;
  if (finalize) {
    c.host_ll = h->d_map;
    c.seq = ++h->seq;
  }
  combine_kernel<<<1, kCombineThreads, 0, h->stream>>>(c);
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  if (finalize) h->cur ^= 1;
  return TAMPC_OK;
}

size_t comm_parity_bytes(int world) { return (sizeof(unsigned long long) * 2 * kPartialWords * world + 255) & ~size_t(255); }

// exchange + merge in one kernel (reduce_kernels.cuh exchange_combine_kernel); this rank's partial is in d_rank_partial
ExchangeArgs make_exchange_args(tampc_mppi* h) {
  ExchangeArgs e{};
  e.c.sampler = h->d_sampler;
  e.c.U = h->d_U[h->cur];
  e.c.T = h->T;
  e.c.nu = h->cfg.nu;
  e.c.finalize = 1;
  e.c.U_out = h->d_U[h->cur ^ 1];
  e.c.action_out = h->d_action;
  e.c.stats_out = h->d_stats;
  e.c.partial_out = h->d_rank_partial;
  e.c.host_ll = h->d_map;
  e.c.seq = ++h->seq;
  e.mine = h->d_rank_partial;
  e.rank = h->comm_rank;
  e.world_size = h->comm_world;
  e.world = h->d_world_partials;
  e.seq = ++h->comm_seq;
  e.timeout_ns = 2000000000ll * (long long)env_int("TAMPC_COMM_TIMEOUT_S", 2) / 2;
  e.error_out = h->d_comm_error;
  const size_t off = (e.seq & 1) * comm_parity_bytes(h->comm_world);
  for (int p = 0; p < h->comm_world; ++p) {
    e.peer_ll[p] = reinterpret_cast<unsigned long long*>(h->comm_peer[p] + off);
  }
  return e;
}

int run_exchange_combine(tampc_mppi* h) {
  const ExchangeArgs e = make_exchange_args(h);
  exchange_combine_kernel<<<1, kCombineThreads, 0, h->stream>>>(e);
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  h->cur ^= 1;
  return TAMPC_OK;
}

int upload_state(tampc_mppi* h, const double* state) {
  memcpy(h->state_host, state, sizeof(double) * h->cfg.nx);  // travels as a kernel argument
  return TAMPC_OK;
}

int upload_noise(tampc_mppi* h, int mode, const double* noise, int T) {
  if (mode == kModePhilox) return TAMPC_OK;
  if (mode != kModeInjected && mode != kModeActions) { h->err = "unknown noise_mode"; return TAMPC_ERR_INVALID; }
  if (!noise) { h->err = "noise_mode needs a host array"; return TAMPC_ERR_INVALID; }
  int rc = ensure_noise(h);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_noise, noise, sizeof(double) * (size_t)h->cfg.num_local * T * h->cfg.nu, cudaMemcpyHostToDevice, h->stream));
  return TAMPC_OK;
}

// The combine kernel stores action and stats straight into mapped pinned memory and then publishes the sequence
// number; poll for it (a few microseconds sooner than a stream synchronise), fall back to the synchronise so that a
// faulted kernel surfaces as an error instead of a hang.
// One value of the result block if both of its words carry the command's sequence number.
bool ll_read(const volatile unsigned long long* base, int slot, uint32_t seq32, double* out) {
  const unsigned long long w0 = base[2 * slot], w1 = base[2 * slot + 1];
  if ((uint32_t)(w0 >> 32) != seq32 || (uint32_t)(w1 >> 32) != seq32) return false;
  const unsigned long long b = (w0 & 0xffffffffull) | (w1 << 32);
  memcpy(out, &b, sizeof(double));
  return true;
}

// 0: not there yet, 1: action + stats complete, -1: the kernel reported a peer-exchange timeout
int ll_poll(tampc_mppi* h, double* vals) {
  const uint32_t seq32 = (uint32_t)h->seq;
  double err;
  if (ll_read(h->h_map, kMaxNu + 3, seq32, &err)) return -1;
  for (int d = 0; d < h->cfg.nu; ++d)
    if (!ll_read(h->h_map, d, seq32, vals + d)) return 0;
  for (int i = 0; i < 3; ++i)
    if (!ll_read(h->h_map, kMaxNu + i, seq32, vals + kMaxNu + i)) return 0;
  return 1;
}

int download_action(tampc_mppi* h, double* action) {
  double vals[kHostSlots];
  int got = 0;
  for (int spin = 0; spin < 200000 && got == 0; ++spin) {
    got = ll_poll(h, vals);
    if (got == 0 && (spin & 1023) == 1023 && cudaStreamQuery(h->stream) != cudaErrorNotReady) break;
  }
  if (got == 0) {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    got = ll_poll(h, vals);
    if (got == 0) { h->err = "combine kernel did not publish its result"; return TAMPC_ERR_CUDA; }
  }
  if (got < 0) {
    // the exchange kernel returned before writing the new nominal sequence: the flip done at launch time would make the
    // next command read a stale buffer, and the ranks' sequence numbers no longer agree — undo the flip, keep the
    // pre-command U, and refuse further sharded commands until the peers are connected again
    h->cur ^= 1;
    h->comm_broken = true;
    cudaMemsetAsync(h->d_comm_error, 0, sizeof(int), h->stream);
    h->err = "peer exchange timed out: a rank did not reach this command (nominal sequence left as before the command; "
             "call comm_disconnect + comm_init + comm_connect on every rank to resume)";
    return TAMPC_ERR_CUDA;
  }
  memcpy(action, vals, sizeof(double) * h->cfg.nu);
  memcpy(h->last_stats, vals + kMaxNu, sizeof(double) * 3);
  return TAMPC_OK;
}

}  // namespace

// ==================================================================================================
extern "C" {

TAMPC_API int tampc_mppi_abi_version(void) { return TAMPC_ABI_VERSION; }

TAMPC_API const char* tampc_mppi_last_error(const tampc_mppi* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

TAMPC_API int tampc_mppi_create(const tampc_mppi_config* cfg, tampc_mppi** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return TAMPC_ERR_INVALID; }
  *out = nullptr;
  if (cfg->abi_version != TAMPC_ABI_VERSION) { g_create_error = "ABI version mismatch"; return TAMPC_ERR_INVALID; }
  if (cfg->nx < 1 || cfg->nx > TAMPC_MAX_NX || cfg->nu < 1 || cfg->nu > TAMPC_MAX_NU) { g_create_error = "nx/nu out of range"; return TAMPC_ERR_INVALID; }
  if (cfg->num_samples < 1 || cfg->num_local < 1 || cfg->sample_offset < 0 || cfg->sample_offset + cfg->num_local > cfg->num_samples) {
    g_create_error = "sample shard [offset, offset+num_local) must lie inside [0, num_samples)";
    return TAMPC_ERR_INVALID;
  }
  if (cfg->max_horizon < 1 || cfg->max_horizon > TAMPC_MAX_HORIZON) { g_create_error = "max_horizon out of range (1..128)"; return TAMPC_ERR_INVALID; }
  if (cfg->precision < TAMPC_PREC_F64 || cfg->precision > TAMPC_PREC_F16X3) { g_create_error = "unknown precision"; return TAMPC_ERR_INVALID; }
  if (!(cfg->lambda_ > 0.0)) { g_create_error = "lambda_ must be positive"; return TAMPC_ERR_INVALID; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    return TAMPC_ERR_NO_DEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return TAMPC_ERR_INVALID; }
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return TAMPC_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "kernels are compiled for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return TAMPC_ERR_NO_DEVICE;
  }
  // Cholesky and inverse of the noise covariance (float64, host)
  const int nu = cfg->nu;
  double L[kMaxNu * kMaxNu] = {0}, inv[kMaxNu * kMaxNu] = {0};
  for (int i = 0; i < nu; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = cfg->noise_sigma[i * nu + j];
      for (int k = 0; k < j; ++k) s -= L[i * kMaxNu + k] * L[j * kMaxNu + k];
      if (i == j) {
        if (!(s > 0.0)) { g_create_error = "noise_sigma is not positive definite"; return TAMPC_ERR_INVALID; }
        L[i * kMaxNu + i] = sqrt(s);
      } else {
        L[i * kMaxNu + j] = s / L[j * kMaxNu + j];
      }
    }
  {  // Sigma^-1 = L^-T L^-1
    double Li[kMaxNu * kMaxNu] = {0};
    for (int c = 0; c < nu; ++c)
      for (int r = 0; r < nu; ++r) {
        if (r < c) continue;
        double s = (r == c) ? 1.0 : 0.0;
        for (int k = c; k < r; ++k) s -= L[r * kMaxNu + k] * Li[k * kMaxNu + c];
        Li[r * kMaxNu + c] = s / L[r * kMaxNu + r];
      }
    for (int i = 0; i < nu; ++i)
      for (int j = 0; j < nu; ++j) {
        double s = 0.0;
        for (int k = 0; k < nu; ++k) s += Li[k * kMaxNu + i] * Li[k * kMaxNu + j];
        inv[i * kMaxNu + j] = s;
      }
  }
  tampc_mppi* h = new tampc_mppi();
  h->cfg = *cfg;
  h->n_sm = prop.multiProcessorCount;
  auto fail = [&](const std::string& msg, int code) {
    g_create_error = msg;
    tampc_mppi_destroy(h);
    return code;
  };
#define CREATE_TRY(expr)                                                                  \
  do {                                                                                    \
    cudaError_t e2 = (expr);                                                              \
    if (e2 != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e2), TAMPC_ERR_CUDA); \
  } while (0)
  if (cfg->stream) {
    h->stream = (cudaStream_t)cfg->stream;
  } else {
    CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  const int Tm = cfg->max_horizon, K = cfg->num_local;
  h->chunk = env_int("TAMPC_REDUCE_CHUNK", K >= (1 << 16) ? 256 : 64);
  if (h->chunk < 1 || h->chunk > kReduceThreads) h->chunk = 64;
  h->n_chunks = (K + h->chunk - 1) / h->chunk;
  CREATE_TRY(cudaMalloc(&h->d_sampler, sizeof(SamplerD)));
  CREATE_TRY(cudaMalloc(&h->d_cost, sizeof(tampc_cost_desc)));
  CREATE_TRY(cudaMalloc(&h->d_U[0], sizeof(double) * Tm * kMaxNu));
  CREATE_TRY(cudaMalloc(&h->d_U[1], sizeof(double) * Tm * kMaxNu));
  CREATE_TRY(cudaMalloc(&h->d_state, sizeof(double) * kMaxNx));
  CREATE_TRY(cudaMalloc(&h->d_costs, sizeof(double) * K));
  h->n_partial_slots = h->n_chunks;
  CREATE_TRY(cudaMalloc(&h->d_partials, sizeof(tampc_partial) * h->n_partial_slots));
  CREATE_TRY(cudaMalloc(&h->d_phdr, sizeof(PartialHdr) * kMaxFusedWarps));
  CREATE_TRY(cudaMalloc(&h->d_pwsum, sizeof(double) * (size_t)kMaxFusedWarps * TAMPC_MAX_HORIZON * TAMPC_MAX_NU));
  CREATE_TRY(cudaMalloc(&h->d_ticket, sizeof(unsigned int)));
  CREATE_TRY(cudaMemset(h->d_ticket, 0, sizeof(unsigned int)));
  h->n_partials2 = (h->n_chunks + kCombineThreads - 1) / kCombineThreads;
  if (h->n_partials2 > kCombineThreads) return fail("num_local too large for the two-level combine", TAMPC_ERR_INVALID);
  CREATE_TRY(cudaMalloc(&h->d_partials2, sizeof(tampc_partial) * h->n_partials2));
  CREATE_TRY(cudaMalloc(&h->d_rank_partial, sizeof(tampc_partial)));
  CREATE_TRY(cudaMalloc(&h->d_action, sizeof(double) * kMaxNu));
  CREATE_TRY(cudaMalloc(&h->d_stats, sizeof(double) * 4));
  CREATE_TRY(cudaMallocHost(&h->h_pin, sizeof(double) * 32));
  CREATE_TRY(cudaHostAlloc(&h->h_map, sizeof(unsigned long long) * 2 * kHostSlots, cudaHostAllocMapped));
  memset(h->h_map, 0, sizeof(unsigned long long) * 2 * kHostSlots);
  CREATE_TRY(cudaHostGetDevicePointer(&h->d_map, h->h_map, 0));
  SamplerD sd{};
  for (int i = 0; i < nu; ++i) {
    sd.mu[i] = cfg->noise_mu[i];
    sd.u_min[i] = cfg->u_min[i];
    sd.u_max[i] = cfg->u_max[i];
    sd.u_init[i] = cfg->u_init[i];
    for (int j = 0; j < nu; ++j) {
      sd.chol[i * kMaxNu + j] = L[i * kMaxNu + j];
      sd.lam_sigma_inv[i * kMaxNu + j] = cfg->lambda_ * inv[i * kMaxNu + j];
    }
  }
  sd.lambda_inv = 1.0 / cfg->lambda_;
  sd.has_bounds = cfg->has_bounds;
  sd.null_action = cfg->sample_null_action;
  CREATE_TRY(cudaMemcpy(h->d_sampler, &sd, sizeof sd, cudaMemcpyHostToDevice));
#undef CREATE_TRY
  *out = h;
  return TAMPC_OK;
}

TAMPC_API void tampc_mppi_destroy(tampc_mppi* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_sampler); cudaFree(h->d_cost); cudaFree(h->d_U[0]); cudaFree(h->d_U[1]); cudaFree(h->d_state);
  cudaFree(h->d_costs); cudaFree(h->d_noise); cudaFree(h->d_states); cudaFree(h->d_scratch); cudaFree(h->d_predict); cudaFree(h->d_select); cudaFree(h->d_tc_w2); cudaFree(h->d_tc_prof); cudaFree(h->d_partials); cudaFree(h->d_partials2); cudaFree(h->d_ticket); cudaFree(h->d_phdr); cudaFree(h->d_pwsum);
  cudaFree(h->d_rank_partial); cudaFree(h->d_action); cudaFree(h->d_stats); cudaFree(h->d_weights);
  for (int p = 0; p < h->comm_world; ++p)
    if (p != h->comm_rank && h->comm_peer[p] && h->comm_peer[p] != h->comm_local) cudaIpcCloseMemHandle(h->comm_peer[p]);
  cudaFree(h->comm_local); cudaFree(h->d_comm_error); cudaFree(h->d_world_partials);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->h_map) cudaFreeHost(h->h_map);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  cudaFree(h->d_gp);
  cudaFree(h->d_gp_eps);
  delete h;
}

}  // extern "C"

// Lay the model out for the kernels the handle will run.  With the GP residual on, the networks are always in the
// FMA layout (gp_rollout.cuh); otherwise the precision picks the layout.
// With a sampled GP and rollout_var_cost > 0 the cost holds the VARIANCE over the M replicas of every step's cost
// (controller.py:364).  The replicas differ only by the GP draws, so that variance is a difference of nearly equal
// numbers, and the independent float32 rounding of each replica's network pass shows up in it amplified ~1000x
// (measured: 6.4e-5 on the total cost with float32 networks and everything else in float64; north star 1e-5).  Such
// a configuration therefore runs the float64 kernel whatever precision was asked for (K is ~500 there: latency-bound).
static bool gp_needs_f64(const tampc_mppi* h) {
  return h->gp_on && h->gp.sample && h->gp.rollout_samples > 1 && h->gp.rollout_var_cost != 0.0;
}

static int install_model(tampc_mppi* h, const tampc_model_desc* desc, const double* params, size_t n_params) {
  PackedModel pm;
  int rc = pack_model(h, *desc, params, n_params, pm);
  if (rc) return rc;
  h->use_tc = false;
  if (pm.tc) {
    if (h->gp_on) { h->err = "the tcgen05 MLP kernel has no GP-residual form"; return TAMPC_ERR_UNSUPPORTED; }
    // image = slabs | CostS<float,float> | SamplerS<float> | SamplerS<double> (the layout of the other tensor-core kernels)
    const size_t wbytes = (pm.tc_image.size() + 15) & ~size_t(15);
    const size_t cost_off = wbytes, samp_off = cost_off + ((sizeof(CostS<float, float>) + 15) & ~size_t(15));
    const size_t samp_d_off = samp_off + ((sizeof(SamplerS<float>) + 15) & ~size_t(15));
    const size_t bytes = samp_d_off + ((sizeof(SamplerS<double>) + 15) & ~size_t(15));
    TcMlpDims d = pm.tc_dims;
    d.off_cost = (int)cost_off;
    d.off_sampler = (int)samp_off;
    // ring stages: what is left of the 227 KB next to the image, the nominal sequence and the control block
    const int ring_off = tc::smem_ring_off((int)bytes, h->cfg.max_horizon);
    const int avail = 226 * 1024 - ring_off;
    int stages = avail / d.w2_stage_bytes;
    if (stages > tc::kMaxStages) stages = tc::kMaxStages;
    if (stages > (tc::kWFullBars - 1) * (d.nc / 8) && stages < d.h1 / 8) stages = (tc::kWFullBars - 1) * (d.nc / 8);  // chunks in flight < barriers
    if (stages >= d.h1 / 8) { stages = d.h1 / 8; d.resident = 1; }
    if (stages < 2 * (d.nc / 8) && !d.resident) { h->err = "tcgen05 MLP kernel: shared memory cannot hold two K-chunks of layer 2"; return TAMPC_ERR_UNSUPPORTED; }
    d.stages = stages;
    d.n_issuers = (d.h2 % 32 == 0 && env_int("TAMPC_TC_ISSUERS", 1) == 2) ? 2 : 1;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (bytes > h->weights_bytes) {
      cudaFree(h->d_weights);
      h->d_weights = nullptr;
      CUDA_TRY(h, cudaMalloc(&h->d_weights, bytes));
      h->weights_bytes = bytes;
    }
    CUDA_TRY(h, cudaMemset(h->d_weights, 0, h->weights_bytes));
    CUDA_TRY(h, cudaMemcpy(h->d_weights, pm.tc_image.data(), pm.tc_image.size(), cudaMemcpyHostToDevice));
    if (pm.tc_w2.size() > h->tc_w2_bytes) {
      cudaFree(h->d_tc_w2); cudaFree(h->d_tc_prof);
      h->d_tc_w2 = nullptr;
      CUDA_TRY(h, cudaMalloc(&h->d_tc_w2, pm.tc_w2.size()));
      h->tc_w2_bytes = pm.tc_w2.size();
    }
    CUDA_TRY(h, cudaMemcpy(h->d_tc_w2, pm.tc_w2.data(), pm.tc_w2.size(), cudaMemcpyHostToDevice));
    if (env_int("TAMPC_TC_PROF", 0) && !h->d_tc_prof) {
      CUDA_TRY(h, cudaMalloc(&h->d_tc_prof, sizeof(long long) * 64));
      CUDA_TRY(h, cudaMemset(h->d_tc_prof, 0, sizeof(long long) * 64));
    }
    h->use_tc = true;
    h->use_mma = true;  // CostS / SamplerS in the float32 tensor-kernel types
    h->img_f64 = false;
    h->tc_dims = d;
    h->tc_fn = pm.tc_fn;
    h->img_cost_off = (int)cost_off;
    h->img_sampler_off = (int)samp_off;
    h->img_sampler_d_off = (int)samp_d_off;
    h->img_bytes = (int)bytes;
    h->n_weight_elems = (int)(wbytes / 4);
    h->variant = pm.variant;
    h->model = *desc;
    h->has_model = true;
    return prep_image(h);
  }
  h->use_generic = false;
  if (pm.generic) {
    if (h->gp_on) { h->err = "the runtime-shaped kernel has no GP-residual form (GP residuals need one of the compiled network shapes)"; return TAMPC_ERR_UNSUPPORTED; }
    // precision -> types: F64 = float64 everywhere; F32 = float32 everywhere; MIXED and the tensor precisions (which have no
    // kernel for an arbitrary shape) = float32 networks, float64 state and cost sum
    const int prec = h->cfg.precision;
    const int types = prec == TAMPC_PREC_F64 ? 0 : (prec == TAMPC_PREC_F32 ? 2 : 1);
    const rollout_launch_fn fn = generic_launcher(h->cfg.nx, h->cfg.nu, types);
    if (!fn) { h->err = "internal: no runtime-shaped kernel for this nx / nu"; return TAMPC_ERR_UNSUPPORTED; }
    const bool f64 = types == 0;
    const void* src = f64 ? (const void*)pm.f64.data() : (const void*)pm.f32.data();
    const size_t raw_bytes = f64 ? 8 * pm.f64.size() : 4 * pm.f32.size();
    const size_t wbytes = (raw_bytes + 15) & ~size_t(15);
    size_t cost_sz, samp_sz;
    if (types == 0) { cost_sz = sizeof(CostS<double, double>); samp_sz = sizeof(SamplerS<double>); }
    else if (types == 1) { cost_sz = sizeof(CostS<double, float>); samp_sz = sizeof(SamplerS<double>); }
    else { cost_sz = sizeof(CostS<float, float>); samp_sz = sizeof(SamplerS<float>); }
    const size_t cost_off = wbytes, samp_off = cost_off + ((cost_sz + 15) & ~size_t(15));
    const size_t bytes = samp_off + ((samp_sz + 15) & ~size_t(15));
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (bytes > h->weights_bytes) {
      cudaFree(h->d_weights);
      h->d_weights = nullptr;
      CUDA_TRY(h, cudaMalloc(&h->d_weights, bytes));
      h->weights_bytes = bytes;
    }
    CUDA_TRY(h, cudaMemset(h->d_weights, 0, h->weights_bytes));
    CUDA_TRY(h, cudaMemcpy(h->d_weights, src, raw_bytes, cudaMemcpyHostToDevice));
    h->gen = f64 ? pm.gen_f64 : pm.gen_f32;
    h->gen.off_cost = (int)cost_off;
    h->gen_fn = fn;
    h->gen_name = types == 0 ? "rollout_generic_kernel f64 (DFMA), run-time layer widths"
                             : (types == 1 ? "rollout_generic_kernel f32 networks (FFMA), f64 state + cost sum, run-time layer widths"
                                           : "rollout_generic_kernel f32 (FFMA), run-time layer widths");
    h->use_generic = true;
    h->use_mma = false;
    h->img_f64 = f64;
    h->n_weight_elems = (int)(wbytes / (f64 ? 8 : 4));
    h->img_cost_off = (int)cost_off;
    h->img_sampler_off = (int)samp_off;
    h->img_sampler_d_off = -1;
    h->img_bytes = (int)bytes;
    h->variant = pm.variant;
    h->model = *desc;
    h->has_model = true;
    return prep_image(h);
  }
  // arithmetic -> layout:  F64 = DMMA layout (FMA layout if the shape has no mma kernel or TAMPC_F64_IMPL=fma),
  // TF32X3 = mma layout, MIXED / F32 = float32 FMA layout
  const int prec = h->cfg.precision;
  const bool f64 = prec == TAMPC_PREC_F64 || gp_needs_f64(h);
  const char* f64_impl = getenv("TAMPC_F64_IMPL");
  bool use_mma = false;
  if (h->gp_on) {
    if (!pm.variant->gp.f64) { h->err = std::string("no GP-residual kernel compiled for ") + pm.variant->name; return TAMPC_ERR_UNSUPPORTED; }
  } else if (prec == TAMPC_PREC_F16X3) {
    use_mma = pm.variant->bytes_mma_f16 != 0;
    if (!use_mma && pm.variant->kind != TAMPC_DYN_PENDULUM) { h->err = std::string("no 3xFP16 tensor-core kernel for ") + pm.variant->name; return TAMPC_ERR_UNSUPPORTED; }
  } else if (prec == TAMPC_PREC_TF32X3) {
    // shapes without network layers (analytic pendulum) have nothing to put on the tensor cores: the float64
    // state/cost arithmetic of TF32X3 is exactly the MIXED kernel there
    use_mma = pm.variant->bytes_mma_tf32 != 0;
    if (!use_mma && pm.variant->kind != TAMPC_DYN_PENDULUM) { h->err = std::string("no tensor-core kernel for ") + pm.variant->name; return TAMPC_ERR_UNSUPPORTED; }
  } else if (f64 && pm.variant->bytes_mma_f64 && !(f64_impl && !strcmp(f64_impl, "fma"))) {
    use_mma = true;
  }
  if (!use_mma && !pm.variant->f64_s1) {
    h->err = std::string(pm.variant->name) + " has tensor-core kernels only: use precision tf32x3" + (pm.variant->bytes_mma_f64 ? " or f64" : "");
    return TAMPC_ERR_UNSUPPORTED;
  }
  const void* src;
  size_t raw_bytes;
  if (use_mma) {
    const bool f16 = prec == TAMPC_PREC_F16X3;
    const std::vector<unsigned char>& blob = f64 ? pm.mma_f64 : (f16 ? pm.mma_f16 : pm.mma_tf32);
    const int expect = f64 ? pm.variant->bytes_mma_f64 : (f16 ? pm.variant->bytes_mma_f16 : pm.variant->bytes_mma_tf32);
    if ((int)blob.size() != expect) {
      h->err = "internal: packed mma size " + std::to_string(blob.size()) + " != kernel layout " + std::to_string(expect) + " for " + pm.variant->name;
      return TAMPC_ERR_INVALID;
    }
    src = blob.data();
    raw_bytes = blob.size();
  } else {
    src = f64 ? (const void*)pm.f64.data() : (const void*)pm.f32.data();
    const size_t elems = f64 ? pm.f64.size() : pm.f32.size();
    const int expect = f64 ? pm.variant->elems_f64 : pm.variant->elems_f32;
    if ((int)elems != expect) {
      h->err = "internal: packed size " + std::to_string(elems) + " != kernel layout " + std::to_string(expect) + " for " + pm.variant->name;
      return TAMPC_ERR_INVALID;
    }
    raw_bytes = (f64 ? 8 : 4) * elems;
  }
  // shared-memory image = packed networks | CostS | SamplerS, each 16-byte aligned (kernel-side layout structs)
  const size_t wbytes = (raw_bytes + 15) & ~size_t(15);
  size_t cost_sz, samp_sz;
  if (f64 || h->gp_on) { cost_sz = sizeof(CostS<double, double>); samp_sz = sizeof(SamplerS<double>); }  // GP kernels: float64 state + cost in every mode
  else if (prec == TAMPC_PREC_MIXED || ((prec == TAMPC_PREC_TF32X3 || prec == TAMPC_PREC_F16X3) && !use_mma)) { cost_sz = sizeof(CostS<double, float>); samp_sz = sizeof(SamplerS<double>); }
  else { cost_sz = sizeof(CostS<float, float>); samp_sz = sizeof(SamplerS<float>); }
  const size_t cost_off = wbytes, samp_off = cost_off + ((cost_sz + 15) & ~size_t(15));
  const size_t samp_d_off = samp_off + ((samp_sz + 15) & ~size_t(15));  // mma kernels: float64 SamplerS copy
  const size_t bytes = samp_d_off + (use_mma ? ((sizeof(SamplerS<double>) + 15) & ~size_t(15)) : 0);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (bytes > h->weights_bytes) {
    cudaFree(h->d_weights);
    h->d_weights = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->d_weights, bytes));
    h->weights_bytes = bytes;
  }
  CUDA_TRY(h, cudaMemset(h->d_weights, 0, h->weights_bytes));
  CUDA_TRY(h, cudaMemcpy(h->d_weights, src, raw_bytes, cudaMemcpyHostToDevice));
  h->n_weight_elems = (int)(wbytes / (f64 ? 8 : 4));
  h->use_mma = use_mma;
  h->img_f64 = f64;
  h->img_cost_off = (int)cost_off;
  h->img_sampler_off = (int)samp_off;
  h->img_sampler_d_off = (int)samp_d_off;
  h->img_bytes = (int)bytes;
  h->variant = pm.variant;
  h->model = *desc;
  h->has_model = true;
  return prep_image(h);
}

extern "C" {

TAMPC_API int tampc_mppi_set_model(tampc_mppi* h, const tampc_model_desc* desc, const double* params, size_t n_params) {
  if (!h || !desc) return TAMPC_ERR_INVALID;
  if (n_params && !params) { h->err = "params is null"; return TAMPC_ERR_INVALID; }
  if (h->gp_on) {  // a new model invalidates the residual fitted around the old one
    h->gp_on = false;
    h->gp = tampc_gp_desc{};
  }
  int rc = install_model(h, desc, params, n_params);
  if (rc) return rc;
  h->model_params.assign(params, params + n_params);
  return TAMPC_OK;
}

}  // extern "C"

// Cholesky factor (lower, in place in A, row-major n x n); false if A is not positive definite
static bool cholesky_lower(std::vector<double>& A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) v -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = v / d;
    }
  }
  return true;
}

template <typename T>
static void pack_gp(std::vector<unsigned char>& blob, const tampc_gp_desc& g, const double* in_scale, const double* in_off, const double* out_mul,
                    const double* X, const std::vector<std::vector<double>>& alpha, const std::vector<std::vector<double>>& Linv) {
  const int n = g.n_points, npad = (n + 3) & ~3, stride = gp_tri_stride(n);
  blob.assign(gp_block_bytes<T>(n, g.n_out), 0);
  GpHeader<T>& H = *reinterpret_cast<GpHeader<T>*>(blob.data());
  H.n = n; H.npad = npad; H.n_in = g.n_in; H.n_out = g.n_out; H.sample = g.sample; H.space = g.space; H.tri_stride = stride;
  for (int i = 0; i < kGpMaxDim; ++i) {
    H.in_scale[i] = i < g.n_in ? (T)in_scale[i] : (T)0;
    H.in_off[i] = i < g.n_in ? (T)in_off[i] : (T)0;
    const bool o = i < g.n_out;
    H.out_mul[i] = o ? (T)out_mul[i] : (T)0;
    H.neg_half_inv_l2[i] = o ? (T)(-0.5 / (g.lengthscale[i] * g.lengthscale[i])) : (T)0;
    H.oscale[i] = o ? (T)g.outputscale[i] : (T)0;
    H.prior_var[i] = o ? (T)(g.outputscale[i] + g.noise[i]) : (T)0;
  }
  T* Xs = reinterpret_cast<T*>(blob.data() + sizeof(GpHeader<T>));
  T* al = Xs + (size_t)npad * kGpMaxDim;
  T* tri = al + (size_t)g.n_out * npad;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < g.n_in; ++i) Xs[(size_t)j * kGpMaxDim + i] = (T)X[(size_t)j * g.n_in + i];
  for (int d = 0; d < g.n_out; ++d) {
    for (int j = 0; j < n; ++j) al[(size_t)d * npad + j] = (T)alpha[d][j];
    T* row = tri + (size_t)d * stride;
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j <= i; ++j) row[j] = (T)Linv[d][(size_t)i * n + j];
      row += (i + 4) & ~3;
    }
  }
}

extern "C" {

TAMPC_API int tampc_mppi_set_gp(tampc_mppi* h, const tampc_gp_desc* g, const double* params, size_t n_params) {
  if (!h || !g) return TAMPC_ERR_INVALID;
  if (!h->has_model) { h->err = "set_model has not been called"; return TAMPC_ERR_STATE; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const bool on = g->n_points > 0;
  if (!on) {
    if (h->gp_on) {
      h->gp_on = false;
      h->gp = tampc_gp_desc{};
      return install_model(h, &h->model, h->model_params.data(), h->model_params.size());
    }
    return TAMPC_OK;
  }
  const int n = g->n_points, nx = h->cfg.nx, nu = h->cfg.nu;
  const int kind = h->model.kind;
  if (kind != TAMPC_DYN_MLP && kind != TAMPC_DYN_REX) { h->err = "a GP residual needs a network dynamics model"; return TAMPC_ERR_UNSUPPORTED; }
  if (n > TAMPC_GP_MAX_POINTS) { h->err = "GP data points exceed TAMPC_GP_MAX_POINTS"; return TAMPC_ERR_UNSUPPORTED; }
  if (g->rollout_samples < 1 || g->rollout_samples > TAMPC_MAX_ROLLOUT_SAMPLES) { h->err = "rollout_samples outside [1, TAMPC_MAX_ROLLOUT_SAMPLES]"; return TAMPC_ERR_UNSUPPORTED; }
  if (g->space != TAMPC_GP_AT_NETWORK && g->space != TAMPC_GP_ORIGINAL) { h->err = "unknown GP space"; return TAMPC_ERR_INVALID; }
  // where the GP sits decides its widths
  const bool latent = kind == TAMPC_DYN_REX && g->space == TAMPC_GP_AT_NETWORK;
  const int want_in = latent ? h->model.net.dims[0] : nx + nu, want_out = latent ? h->model.nv : nx;
  if (g->n_in != want_in || g->n_out != want_out || g->n_in > kGpMaxDim || g->n_out > kGpMaxDim) {
    h->err = "GP input/output widths do not match the model (" + std::to_string(g->n_in) + "->" + std::to_string(g->n_out) + ", expected " +
             std::to_string(want_in) + "->" + std::to_string(want_out) + ")";
    return TAMPC_ERR_INVALID;
  }
  if (!params || g->x_off < 0 || g->resid_off < 0 || (size_t)g->x_off + (size_t)n * g->n_in > n_params || (size_t)g->resid_off + (size_t)n * g->n_out > n_params) {
    h->err = "GP data offsets outside the params array";
    return TAMPC_ERR_INVALID;
  }
  for (int d = 0; d < g->n_out; ++d)
    if (!(g->lengthscale[d] > 0.0 && g->outputscale[d] > 0.0 && g->noise[d] > 0.0)) { h->err = "GP hyper-parameters must be positive"; return TAMPC_ERR_INVALID; }
  // scalers between the model's arrays and the GP's space
  double in_scale[kGpMaxDim], in_off[kGpMaxDim], out_mul[kGpMaxDim];
  for (int i = 0; i < kGpMaxDim; ++i) { in_scale[i] = 1.0; in_off[i] = 0.0; out_mul[i] = 1.0; }
  if (g->space == TAMPC_GP_ORIGINAL) {
    for (int i = 0; i < g->n_in; ++i) { in_scale[i] = g->in_scale[i]; in_off[i] = g->in_offset[i]; }
    for (int i = 0; i < g->n_out; ++i) {
      if (g->out_scale[i] == 0.0) { h->err = "GP out_scale must be non-zero"; return TAMPC_ERR_INVALID; }
      out_mul[i] = 1.0 / g->out_scale[i];
    }
  } else if (kind == TAMPC_DYN_MLP) {  // the network's own in / out scalers (folded into its layers for the nominal part)
    for (int i = 0; i < g->n_in; ++i) { in_scale[i] = h->net_in_scale[i]; in_off[i] = h->net_in_off[i]; }
    for (int i = 0; i < g->n_out; ++i) out_mul[i] = 1.0 / h->net_out_scale[i];
  } else {  // REX latent: the encoder already emits scaled z; v is un-scaled by the latent net's folded last layer
    for (int i = 0; i < g->n_out; ++i) out_mul[i] = 1.0 / h->net_out_scale[i];
  }
  // exact-GP factors in float64: K_d + n_d I = L L^T, alpha_d = (K_d + n_d I)^-1 resid_d, L^-1
  const double* X = params + g->x_off;
  const double* R = params + g->resid_off;
  std::vector<double> d2((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int c = 0; c < g->n_in; ++c) { const double d = X[(size_t)i * g->n_in + c] - X[(size_t)j * g->n_in + c]; s += d * d; }
      d2[(size_t)i * n + j] = s;
    }
  std::vector<std::vector<double>> alpha(g->n_out), Linv(g->n_out);
  for (int d = 0; d < g->n_out; ++d) {
    std::vector<double> Lm((size_t)n * n);
    const double l2 = g->lengthscale[d] * g->lengthscale[d];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) Lm[(size_t)i * n + j] = g->outputscale[d] * std::exp(-0.5 * d2[(size_t)i * n + j] / l2) + (i == j ? g->noise[d] : 0.0);
    if (!cholesky_lower(Lm, n)) { h->err = "GP kernel matrix of output " + std::to_string(d) + " is not positive definite"; return TAMPC_ERR_INVALID; }
    std::vector<double>& Li = Linv[d];
    Li.assign((size_t)n * n, 0.0);
    for (int c = 0; c < n; ++c) {  // column c of L^-1 by forward substitution on e_c
      for (int i = c; i < n; ++i) {
        double v = i == c ? 1.0 : 0.0;
        for (int k = c; k < i; ++k) v -= Lm[(size_t)i * n + k] * Li[(size_t)k * n + c];
        Li[(size_t)i * n + c] = v / Lm[(size_t)i * n + i];
      }
    }
    std::vector<double> w(n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) w[i] += Li[(size_t)i * n + j] * R[(size_t)j * g->n_out + d];
    alpha[d].assign(n, 0.0);
    for (int j = 0; j < n; ++j)
      for (int i = j; i < n; ++i) alpha[d][j] += Li[(size_t)i * n + j] * w[i];
  }
  std::vector<unsigned char> blob;
  pack_gp<double>(blob, *g, in_scale, in_off, out_mul, X, alpha, Linv);  // the GP part is float64 in every precision mode (gp_rollout.cuh)
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (blob.size() > h->gp_capacity) {
    cudaFree(h->d_gp);
    h->d_gp = nullptr;
    const size_t cap = gp_block_bytes<double>(TAMPC_GP_MAX_POINTS, kGpMaxDim);
    CUDA_TRY(h, cudaMalloc(&h->d_gp, cap));
    h->gp_capacity = cap;
  }
  CUDA_TRY(h, cudaMemcpy(h->d_gp, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  h->gp_bytes = (int)blob.size();
  h->gp = *g;
  const bool was_on = h->gp_on;
  h->gp_on = true;
  if (!was_on || (h->img_f64 != (h->cfg.precision == TAMPC_PREC_F64 || gp_needs_f64(h)))) {
    int rc = install_model(h, &h->model, h->model_params.data(), h->model_params.size());
    if (rc) { h->gp_on = was_on; return rc; }
  }
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_set_gp_draws(tampc_mppi* h, const double* eps, size_t n) {
  if (!h) return TAMPC_ERR_INVALID;
  if (!eps) { h->gp_eps_armed = false; return TAMPC_OK; }
  if (!h->gp_on) { h->err = "set_gp has not been called"; return TAMPC_ERR_STATE; }
  const size_t per_step = (size_t)h->cfg.num_local * h->gp.rollout_samples * h->gp.n_out;
  if (n == 0 || n % per_step != 0 || n / per_step > (size_t)h->cfg.max_horizon) { h->err = "GP draws must be (num_local, M, T, n_out)"; return TAMPC_ERR_INVALID; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (n > h->gp_eps_capacity) {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_gp_eps);
    h->d_gp_eps = nullptr;
    CUDA_TRY(h, cudaMalloc(&h->d_gp_eps, sizeof(double) * n));
    h->gp_eps_capacity = n;
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->d_gp_eps, eps, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->gp_eps_armed = true;
  h->gp_eps_T = (int)(n / per_step);
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_set_cost(tampc_mppi* h, const tampc_cost_desc* d) {
  if (!h || !d) return TAMPC_ERR_INVALID;
  if (d->nx != h->cfg.nx || d->nu != h->cfg.nu) { h->err = "cost nx/nu differ from the handle's"; return TAMPC_ERR_INVALID; }
  if (d->kind == TAMPC_COST_GOAL_TRAP) {
    if (d->n_traps < 0 || d->n_traps > TAMPC_MAX_TRAPS) { h->err = "n_traps out of range"; return TAMPC_ERR_INVALID; }
  } else if (d->kind == TAMPC_COST_RECOVERY) {
    if (d->n_sets < 1 || d->n_sets > TAMPC_MAX_GOAL_SETS) { h->err = "n_sets out of range"; return TAMPC_ERR_INVALID; }
    for (int j = 0; j < d->n_sets; ++j)
      if (d->set_size[j] < 1 || d->set_size[j] > TAMPC_MAX_SET_GOALS) { h->err = "goal set size out of range"; return TAMPC_ERR_INVALID; }
  } else if (d->kind == TAMPC_COST_APF) {
    if (d->n_traps < 0 || d->n_traps > TAMPC_MAX_TRAPS) { h->err = "n_traps out of range"; return TAMPC_ERR_INVALID; }
    if (!(d->apf_max_dist > 0.0)) { h->err = "apf_max_dist must be positive"; return TAMPC_ERR_INVALID; }
  } else {
    h->err = "unknown cost kind";
    return TAMPC_ERR_INVALID;
  }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  // stream-ordered with respect to earlier commands; the pageable source is consumed before the call returns
  CUDA_TRY(h, cudaMemcpyAsync(h->d_cost, d, sizeof *d, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->has_cost = true;
  return prep_image(h);
}

TAMPC_API int tampc_mppi_set_nominal(tampc_mppi* h, const double* U, int32_t horizon) {
  if (!h || !U) return TAMPC_ERR_INVALID;
  if (horizon < 1 || horizon > h->cfg.max_horizon) { h->err = "horizon outside [1, max_horizon]"; return TAMPC_ERR_INVALID; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_U[h->cur], U, sizeof(double) * horizon * h->cfg.nu, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->T = horizon;
  h->has_nominal = true;
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_get_nominal(tampc_mppi* h, double* U) {
  if (!h || !U) return TAMPC_ERR_INVALID;
  if (!h->has_nominal) { h->err = "set_nominal has not been called"; return TAMPC_ERR_STATE; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaMemcpyAsync(U, h->d_U[h->cur], sizeof(double) * h->T * h->cfg.nu, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_command_begin(tampc_mppi* h, const double* state, int32_t noise_mode, const double* noise, int32_t keep_states,
                             void* partial_dev) {
  if (!h || !state) return TAMPC_ERR_INVALID;
  int rc = check_ready(h);
  if (rc) return rc;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if ((rc = upload_state(h, state))) return rc;
  if ((rc = upload_noise(h, noise_mode, noise, h->T))) return rc;
  if (keep_states && (rc = ensure_states(h))) return rc;
  if (!partial_dev) { h->err = "partial_dev is null"; return TAMPC_ERR_INVALID; }
  bool done = false;
  if ((rc = run_rollout_and_partials(h, noise_mode, keep_states != 0, /*fuse=*/2, (tampc_partial*)partial_dev, &done))) return rc;
  if (done) return TAMPC_OK;
  return run_combine(h, h->d_partials, h->n_chunks, /*finalize=*/0, (tampc_partial*)partial_dev);
}

TAMPC_API int tampc_mppi_command_finish(tampc_mppi* h, const void* partials_dev, int32_t world, double* action) {
  if (!h || !partials_dev || !action || world < 1) return TAMPC_ERR_INVALID;
  int rc = check_ready(h);
  if (rc) return rc;
  if (!h->last_valid) { h->err = "command_finish without command_begin"; return TAMPC_ERR_STATE; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if ((rc = run_combine(h, (const tampc_partial*)partials_dev, world, /*finalize=*/1))) return rc;
  return download_action(h, action);
}

TAMPC_API int tampc_mppi_command(tampc_mppi* h, const double* state, int32_t noise_mode, const double* noise, double* action) {
  if (!h || !state || !action) return TAMPC_ERR_INVALID;
  int rc = check_ready(h);
  if (rc) return rc;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if ((rc = upload_state(h, state))) return rc;
  if ((rc = upload_noise(h, noise_mode, noise, h->T))) return rc;
  bool done = false;
  if (h->comm_connected && h->comm_broken) { h->err = "peer exchange broken by an earlier timeout: reconnect the ranks first"; return TAMPC_ERR_STATE; }
  if (h->comm_connected) {
    // K-sharded: rank partial (fused into the rollout or via the partial kernels), then exchange + merge in one kernel
    if ((rc = run_rollout_and_partials(h, noise_mode, false, /*fuse=*/3, h->d_rank_partial, &done))) return rc;
    if (!done && (rc = run_combine(h, h->d_partials, h->n_chunks, /*finalize=*/0, h->d_rank_partial))) return rc;
    if (!h->tail_exchanged && (rc = run_exchange_combine(h))) return rc;
    return download_action(h, action);  // (a peer-exchange timeout comes back through the result block's error slot)
  }
  if (h->cfg.num_local != h->cfg.num_samples) {
    h->err = "sharded handle: use command_begin/finish with an all-gather, or connect the peer exchange (comm_init/comm_connect)";
    return TAMPC_ERR_STATE;
  }
  if ((rc = run_rollout_and_partials(h, noise_mode, false, /*fuse=*/1, nullptr, &done))) return rc;
  if (!done && (rc = run_combine(h, h->d_partials, h->n_chunks, /*finalize=*/1))) return rc;
  return download_action(h, action);
}

TAMPC_API int tampc_mppi_comm_init(tampc_mppi* h, int32_t rank, int32_t world, void* handle_out) {
  if (!h || !handle_out) return TAMPC_ERR_INVALID;
  if (world < 2 || world > kMaxRanks || rank < 0 || rank >= world) { h->err = "comm_init: need 2 <= world <= 8 and 0 <= rank < world"; return TAMPC_ERR_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == TAMPC_IPC_HANDLE_BYTES, "IPC handle size");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (h->comm_connected) { h->err = "comm_init on a connected handle: call comm_disconnect first"; return TAMPC_ERR_STATE; }
  {
    // (re)allocate for this world size; zeroed in stream order with the kernels that will poll it, and the zeroes are in
    // memory before the IPC handle leaves this process (a peer's first store can only follow its comm_connect)
    const size_t bytes = 2 * comm_parity_bytes(world);
    if (h->comm_local && world != h->comm_world) { cudaFree(h->comm_local); h->comm_local = nullptr; }
    if (!h->comm_local) CUDA_TRY(h, cudaMalloc(&h->comm_local, bytes));
    if (!h->d_comm_error) CUDA_TRY(h, cudaMalloc(&h->d_comm_error, sizeof(int)));
    if (!h->d_world_partials) CUDA_TRY(h, cudaMalloc(&h->d_world_partials, sizeof(tampc_partial) * kMaxRanks));
    CUDA_TRY(h, cudaMemsetAsync(h->comm_local, 0, bytes, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->d_comm_error, 0, sizeof(int), h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->comm_seq = 0;
    h->comm_broken = false;
  }
  h->comm_rank = rank;
  h->comm_world = world;
  cudaIpcMemHandle_t hd;
  CUDA_TRY(h, cudaIpcGetMemHandle(&hd, h->comm_local));
  memcpy(handle_out, &hd, sizeof hd);
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_comm_connect(tampc_mppi* h, const void* handles) {
  if (!h || !handles) return TAMPC_ERR_INVALID;
  if (!h->comm_local) { h->err = "comm_connect before comm_init"; return TAMPC_ERR_STATE; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const unsigned char* hb = reinterpret_cast<const unsigned char*>(handles);
  for (int p = 0; p < h->comm_world; ++p) {
    if (p == h->comm_rank) { h->comm_peer[p] = h->comm_local; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, hb + (size_t)p * TAMPC_IPC_HANDLE_BYTES, sizeof hd);
    void* ptr = nullptr;
    const cudaError_t oe = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
    if (oe != cudaSuccess) {
      for (int q = 0; q < p; ++q) {
        if (q != h->comm_rank && h->comm_peer[q]) cudaIpcCloseMemHandle(h->comm_peer[q]);
        h->comm_peer[q] = nullptr;
      }
      h->comm_peer[h->comm_rank] = nullptr;
      h->err = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(p) + "): " + cudaGetErrorString(oe);
      return TAMPC_ERR_CUDA;
    }
    h->comm_peer[p] = reinterpret_cast<unsigned char*>(ptr);
  }
  h->comm_connected = true;
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_comm_disconnect(tampc_mppi* h) {
  if (!h) return TAMPC_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (int p = 0; p < h->comm_world; ++p) {
    if (p != h->comm_rank && h->comm_peer[p] && h->comm_peer[p] != h->comm_local) cudaIpcCloseMemHandle(h->comm_peer[p]);
    h->comm_peer[p] = nullptr;
  }
  h->comm_connected = false;
  h->comm_broken = false;
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_set_state_resident(tampc_mppi* h, const double* state) {
  if (!h || !state) return TAMPC_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  int rc = upload_state(h, state);
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_command_resident(tampc_mppi* h, int32_t noise_mode) {
  if (!h) return TAMPC_ERR_INVALID;
  int rc = check_ready(h);
  if (rc) return rc;
  if (noise_mode != kModePhilox && !h->d_noise) { h->err = "resident command with injected noise needs a previous upload"; return TAMPC_ERR_STATE; }
  bool done = false;
  if (h->comm_connected && h->comm_broken) { h->err = "peer exchange broken by an earlier timeout: reconnect the ranks first"; return TAMPC_ERR_STATE; }
  if (h->comm_connected) {
    if ((rc = run_rollout_and_partials(h, noise_mode, false, /*fuse=*/3, h->d_rank_partial, &done))) return rc;
    if (!done && (rc = run_combine(h, h->d_partials, h->n_chunks, /*finalize=*/0, h->d_rank_partial))) return rc;
    return h->tail_exchanged ? TAMPC_OK : run_exchange_combine(h);
  }
  if ((rc = run_rollout_and_partials(h, noise_mode, false, /*fuse=*/1, nullptr, &done))) return rc;
  if (done) return TAMPC_OK;
  return run_combine(h, h->d_partials, h->n_chunks, /*finalize=*/1);
}

TAMPC_API int tampc_mppi_rollout_resident(tampc_mppi* h, int32_t noise_mode) {
  if (!h) return TAMPC_ERR_INVALID;
  int rc = check_ready(h);
  if (rc) return rc;
  if (noise_mode != kModePhilox && !h->d_noise) { h->err = "resident rollout with injected noise needs a previous upload"; return TAMPC_ERR_STATE; }
  return launch_rollout_for(h, noise_mode, /*shift=*/1, h->d_U[h->cur], h->T, h->call + 1, false);
}

TAMPC_API int tampc_mppi_rollout_costs(tampc_mppi* h, const double* state, const double* actions, int32_t horizon, double* cost_out, double* states_out) {
  if (!h || !state || !actions || !cost_out) return TAMPC_ERR_INVALID;
  if (!h->has_model || !h->has_cost) { h->err = "set_model / set_cost first"; return TAMPC_ERR_STATE; }
  if (horizon < 1 || horizon > h->cfg.max_horizon) { h->err = "horizon outside [1, max_horizon]"; return TAMPC_ERR_INVALID; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  int rc;
  if ((rc = upload_state(h, state))) return rc;
  if ((rc = upload_noise(h, kModeActions, actions, horizon))) return rc;
  if (states_out && (rc = ensure_states(h))) return rc;
  if ((rc = launch_rollout_for(h, kModeActions, 0, h->d_U[h->cur], horizon, 0, states_out != nullptr))) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(cost_out, h->d_costs, sizeof(double) * h->cfg.num_local, cudaMemcpyDeviceToHost, h->stream));
  if (states_out)
    CUDA_TRY(h, cudaMemcpyAsync(states_out, h->d_states, sizeof(double) * (size_t)h->cfg.num_local * horizon * h->cfg.nx, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->last_valid = false;
  h->states_valid = false;
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_get_rollouts(tampc_mppi* h, const double* state, double* states_out) {
  if (!h || !state || !states_out) return TAMPC_ERR_INVALID;
  int rc = check_ready(h);
  if (rc) return rc;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if ((rc = upload_state(h, state))) return rc;
  if ((rc = ensure_states(h))) return rc;
  // one sample, nominal actions, states dumped; the per-sample cost buffer is left untouched via a K_local=1 view
  RolloutArgs a{};
  a.image = h->d_weights;
  a.image_bytes = h->img_bytes;
  a.tc = h->tc_dims;
  a.gen = h->gen;
  a.tc_w2 = h->d_tc_w2;
  a.tc_prof = h->d_tc_prof;
  a.slope = h->model.leaky_slope;
  a.cost = h->d_cost;
  a.sampler = h->d_sampler;
  a.U = h->d_U[h->cur];
  a.shift = 0;
  memcpy(a.state, h->state_host, sizeof a.state);
  a.src = make_src(h, kModeNominal, 0, h->T);
  a.K_local = 1;
  a.cost_out = h->d_stats + 3;
  a.states_out = h->d_states;
  a.wrap_mask = h->model.wrap_mask;
  a.has_guard = h->model.has_bad_state_guard;
  a.bad_norm = h->model.bad_state_norm;
  if ((rc = apply_gp_args(h, a, h->T))) return rc;
  const LaunchPlan plan = choose_launch(h, 1);
  if (!plan.fn) { h->err = "no kernel compiled for this precision"; return TAMPC_ERR_UNSUPPORTED; }
  CUDA_TRY(h, plan.fn(a, plan.block, h->stream));
  h->launches++;
  CUDA_TRY(h, cudaMemcpyAsync(states_out, h->d_states, sizeof(double) * h->T * h->cfg.nx, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->states_valid = false;
  return TAMPC_OK;
}

// One step from `state` under `action`.  nominal_only = false: OnlineMPC._apply_dynamics (guard, network + GP residual,
// constrain_state; controller.py:312-313).  nominal_only = true: OnlineMPC.predict_next_state (online_controller.py:47-58),
// which swaps in the ORIGINAL nominal network for the call — no GP residual, no guard, no constrain_state.
static int predict_impl(tampc_mppi* h, const double* state, const double* action, double* next_state, bool nominal_only) {
  if (!h || !state || !action || !next_state) return TAMPC_ERR_INVALID;
  if (!h->has_model || !h->has_cost) { h->err = "set_model / set_cost first"; return TAMPC_ERR_STATE; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  int rc;
  if ((rc = upload_state(h, state))) return rc;
  if ((rc = ensure_states(h))) return rc;
  // a one-sample, one-step "nominal" rollout whose nominal sequence is the given action
  if (!h->d_predict) CUDA_TRY(h, cudaMalloc(&h->d_predict, sizeof(double) * (kMaxNu + 1)));
  double* d_act = h->d_predict;
  CUDA_TRY(h, cudaMemcpyAsync(d_act, action, sizeof(double) * h->cfg.nu, cudaMemcpyHostToDevice, h->stream));
  RolloutArgs a{};
  a.image = h->d_weights;
  a.image_bytes = h->img_bytes;
  a.tc = h->tc_dims;
  a.gen = h->gen;
  a.tc_w2 = h->d_tc_w2;
  a.tc_prof = h->d_tc_prof;
  a.slope = h->model.leaky_slope;
  a.cost = h->d_cost;
  a.sampler = h->d_sampler;
  a.U = d_act;
  a.shift = 0;
  memcpy(a.state, h->state_host, sizeof a.state);
  a.src = make_src(h, kModeNominal, 0, 1);
  a.K_local = 1;
  a.cost_out = h->d_predict + kMaxNu;
  a.states_out = h->d_states;
  a.wrap_mask = nominal_only ? 0u : h->model.wrap_mask;
  a.has_guard = nominal_only ? 0 : h->model.has_bad_state_guard;
  a.bad_norm = h->model.bad_state_norm;
  LaunchPlan plan;
  if (nominal_only) {
    // with a GP installed the image is in the FMA layout of the (sample, replica) kernels: the plain FMA kernels of
    // the same precision read it as it is and know nothing of the residual
    if (h->gp_on) {
      plan.fn = h->img_f64 ? h->variant->f64_s1 : h->variant->gp.nominal;
      plan.block = 32;
      plan.name = "rollout_kernel (nominal network only)";
    } else {
      plan = choose_launch(h, 1);
    }
  } else {
    if ((rc = apply_gp_args(h, a, 1))) return rc;
    plan = choose_launch(h, 1);
  }
  if (!plan.fn) { h->err = "no kernel compiled for this precision"; return TAMPC_ERR_UNSUPPORTED; }
  CUDA_TRY(h, plan.fn(a, plan.block, h->stream));
  h->launches++;
  CUDA_TRY(h, cudaMemcpyAsync(next_state, h->d_states, sizeof(double) * h->cfg.nx, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->states_valid = false;
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_predict(tampc_mppi* h, const double* state, const double* action, double* next_state) {
  return predict_impl(h, state, action, next_state, false);
}

TAMPC_API int tampc_mppi_predict_nominal(tampc_mppi* h, const double* state, const double* action, double* next_state) {
  return predict_impl(h, state, action, next_state, true);
}

TAMPC_API int tampc_mppi_query(tampc_mppi* h, int32_t what, double* out, size_t n) {
  if (!h || !out) return TAMPC_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const size_t K = h->cfg.num_local;
  switch (what) {
    case TAMPC_Q_COST_TOTAL: {
      if (n < K) { h->err = "query buffer too small"; return TAMPC_ERR_INVALID; }
      CUDA_TRY(h, cudaMemcpyAsync(out, h->d_costs, sizeof(double) * K, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      return TAMPC_OK;
    }
    case TAMPC_Q_OMEGA: {
      if (!h->last_valid) { h->err = "no command has run"; return TAMPC_ERR_STATE; }
      if (n < K) { h->err = "query buffer too small"; return TAMPC_ERR_INVALID; }
      CUDA_TRY(h, cudaMemcpyAsync(out, h->d_costs, sizeof(double) * K, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      const double li = 1.0 / h->cfg.lambda_;
      for (size_t k = 0; k < K; ++k) out[k] = exp(-li * (out[k] - h->last_stats[0])) / h->last_stats[1];
      return TAMPC_OK;
    }
    case TAMPC_Q_NOMINAL: {
      if (n < (size_t)h->T * h->cfg.nu) { h->err = "query buffer too small"; return TAMPC_ERR_INVALID; }
      return tampc_mppi_get_nominal(h, out);
    }
    case TAMPC_Q_STATS: {
      if (n < 4) { h->err = "query buffer too small"; return TAMPC_ERR_INVALID; }
      out[0] = h->last_stats[0]; out[1] = h->last_stats[1]; out[2] = h->last_stats[2]; out[3] = (double)h->launches;
      return TAMPC_OK;
    }
    case TAMPC_Q_NOISE:
    case TAMPC_Q_PERTURBED: {
      if (!h->last_valid) { h->err = "no command has run"; return TAMPC_ERR_STATE; }
      const size_t total = K * h->last_T * h->cfg.nu;
      if (n < total) { h->err = "query buffer too small"; return TAMPC_ERR_INVALID; }
      if (!h->d_scratch) CUDA_TRY(h, cudaMalloc(&h->d_scratch, sizeof(double) * K * h->cfg.max_horizon * h->cfg.nu));
      PerturbedArgs p{};
      p.sampler = h->d_sampler;
      p.U = h->d_U[h->last_U];
      p.src = make_src(h, h->last_mode, h->last_call, h->last_T);
      p.K_local = (int)K;
      p.out = h->d_scratch;
      p.post_clamp_noise = what == TAMPC_Q_NOISE;
      const long long items = (long long)K * h->last_T;
      const int grid = (int)((items + 255) / 256);
      switch (h->cfg.nu) {
        case 1: perturbed_kernel<1><<<grid, 256, 0, h->stream>>>(p); break;
        case 2: perturbed_kernel<2><<<grid, 256, 0, h->stream>>>(p); break;
        case 3: perturbed_kernel<3><<<grid, 256, 0, h->stream>>>(p); break;
        default: perturbed_kernel<4><<<grid, 256, 0, h->stream>>>(p); break;
      }
      CUDA_TRY(h, cudaGetLastError());
      h->launches++;
      CUDA_TRY(h, cudaMemcpyAsync(out, h->d_scratch, sizeof(double) * total, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      return TAMPC_OK;
    }
    case TAMPC_Q_STATES: {
      if (!h->states_valid) { h->err = "last command did not keep states"; return TAMPC_ERR_STATE; }
      const size_t total = K * h->last_T * h->cfg.nx;
      if (n < total) { h->err = "query buffer too small"; return TAMPC_ERR_INVALID; }
      CUDA_TRY(h, cudaMemcpyAsync(out, h->d_states, sizeof(double) * total, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      return TAMPC_OK;
    }
    default: h->err = "unknown query"; return TAMPC_ERR_INVALID;
  }
}

TAMPC_API int tampc_mppi_select_action(tampc_mppi* h, const double* state, const double* actions, int32_t n, int64_t* index_out, double* cost_out,
                                       double* next_state_out) {
  if (!h || !state || !actions || !index_out) return TAMPC_ERR_INVALID;
  if (!h->has_model || !h->has_cost) { h->err = "set_model / set_cost first"; return TAMPC_ERR_STATE; }
  if (n < 1 || n > h->cfg.num_local) { h->err = "select_action: need 1 <= n <= num_local candidate actions"; return TAMPC_ERR_INVALID; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  int rc;
  if ((rc = upload_state(h, state))) return rc;
  if ((rc = ensure_noise(h)) || (rc = ensure_states(h))) return rc;
  const int nx = h->cfg.nx, nu = h->cfg.nu;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_noise, actions, sizeof(double) * (size_t)n * nu, cudaMemcpyHostToDevice, h->stream));
  // a one-step rollout of the n given actions with the states kept; only the first n rows exist for the kernels
  const int saved_local = h->cfg.num_local;
  h->cfg.num_local = n;
  rc = launch_rollout_for(h, kModeActions, 0, h->d_U[h->cur], 1, 0, /*keep_states=*/true);
  h->cfg.num_local = saved_local;
  if (rc) return rc;
  if (!h->d_select) CUDA_TRY(h, cudaMalloc(&h->d_select, sizeof(double) * (2 + kMaxNx)));
  argmin_select_kernel<<<1, 1024, 0, h->stream>>>(h->d_costs, h->d_states, n, nx, h->d_select);
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  double out[2 + kMaxNx];
  CUDA_TRY(h, cudaMemcpyAsync(out, h->d_select, sizeof(double) * (2 + nx), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  *index_out = (int64_t)out[0];
  if (cost_out) *cost_out = out[1];
  if (next_state_out) memcpy(next_state_out, out + 2, sizeof(double) * nx);
  h->last_valid = false;
  h->states_valid = false;
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_state_costs(tampc_mppi* h, const double* states, int32_t n, double* goal_cost_out, double* trap_cost_out) {
  if (!h || !states || !goal_cost_out || !trap_cost_out || n < 1) return TAMPC_ERR_INVALID;
  if (!h->has_cost) { h->err = "set_cost first"; return TAMPC_ERR_STATE; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const int nx = h->cfg.nx;
  double* buf = nullptr;
  CUDA_TRY(h, cudaMallocAsync(&buf, sizeof(double) * (size_t)n * (nx + 2), h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(buf, states, sizeof(double) * (size_t)n * nx, cudaMemcpyHostToDevice, h->stream));
  state_costs_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_cost, buf, n, buf + (size_t)n * nx, buf + (size_t)n * (nx + 1));
  cudaError_t le = cudaGetLastError();
  if (le == cudaSuccess) le = cudaMemcpyAsync(goal_cost_out, buf + (size_t)n * nx, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
  if (le == cudaSuccess) le = cudaMemcpyAsync(trap_cost_out, buf + (size_t)n * (nx + 1), sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
  cudaFreeAsync(buf, h->stream);
  CUDA_TRY(h, le);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->launches++;
  return TAMPC_OK;
}

// development: the 64 clock stamps the tcgen05 kernel left (TAMPC_TC_PROF=1 at set_model time); zeros otherwise
TAMPC_API int tampc_mppi_debug_stamps(tampc_mppi* h, int64_t* out) {
  if (!h || !out) return TAMPC_ERR_INVALID;
  memset(out, 0, sizeof(int64_t) * 64);
  if (!h->d_tc_prof) return TAMPC_OK;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMemcpy(out, h->d_tc_prof, sizeof(int64_t) * 64, cudaMemcpyDeviceToHost));
  return TAMPC_OK;
}

TAMPC_API const char* tampc_mppi_kernel_name(tampc_mppi* h) {
  if (!h || !h->has_model) return "none";
  return choose_launch(h, h->cfg.num_local).name;
}

TAMPC_API int tampc_mppi_device_costs(tampc_mppi* h, void** dev_ptr) {
  if (!h || !dev_ptr) return TAMPC_ERR_INVALID;
  *dev_ptr = h->d_costs;
  return TAMPC_OK;
}

TAMPC_API int tampc_mppi_synchronize(tampc_mppi* h) {
  if (!h) return TAMPC_ERR_INVALID;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return TAMPC_OK;
}

}  // extern "C"
