// variants.cuh — the compiled dynamics shapes.  Each lives in its own translation unit (parallel nvcc).
#pragma once
#include <initializer_list>
#include "mma_rollout.cuh"
#include "gp_rollout.cuh"

namespace tampc {

// One dynamics shape = the set of kernels compiled for it.
struct ModelVariant {
  const char* name;
  int kind, nx, nu;
  int dims[12];            // kind-specific layer widths, see pack_model() in mppi_api.cu
  int elems_f32, elems_f64;  // packed sizes of the FMA-path layouts (elements)
  int bytes_mma_f64, bytes_mma_tf32;  // packed sizes of the mma-path layouts (bytes); 0 = not compiled
  int macs;
  rollout_launch_fn f64_s1, mixed_s1, mixed_s2, f32_s1, f32_s2;  // one-thread-per-sample FMA kernels
  rollout_launch_fn f64_mma[3];   // DMMA kernels, NM = 1, 2, 4 m-tiles (8, 16, 32 samples) per warp
  rollout_launch_fn tf32_mma[2];  // 3xTF32 kernels, NM = 1, 2 (16, 32 samples per warp); float32 state and cost terms, float64 cost sum
  GpLaunchers gp;                 // one-thread-per-(sample, replica) FMA kernels with the online GP residual
  rollout_launch_fn tf32_mma_wide; // NM = 2, 256-thread CTAs, two per SM (throughput-bound K)
  // 3xFP16 (mma.sync.m16n8k16.f16): same kernels, half the tensor instructions
  int bytes_mma_f16;
  rollout_launch_fn f16_mma[2], f16_mma_wide, f16_split;
  rollout_launch_fn tf32_split3, f16_split3;  // three warps per tile (chain | decoder | cost): latency-bound K
  rollout_launch_fn tf32_mma_sm, f16_mma_sm;  // NM = 2, ONE CTA per SM of up to 16 warps: a single balanced wave (K just above 56k)
  rollout_launch_fn tf32_mma_mix, f16_mma_mix;  // one CTA per SM, seven tiles per sub-partition (56 832 < K <= 66 304)
  rollout_launch_fn tf32_mma_persist, f16_mma_persist;  // one 16-warp CTA per SM walking the unit list slot-major (several balanced waves)
  rollout_launch_fn tf32_split;   // two warps per 16-sample tile, split by role (RexExtract shapes, latency-bound K)
  rollout_launch_fn tf32_pipe;    // two-warp pipelined 3xTF32 kernel for small K (16 samples per CTA)
};

template <class MF, class MD, bool WITH_S2>
ModelVariant make_variant(const char* name, int kind, std::initializer_list<int> dims) {
  ModelVariant v{};
  v.name = name;
  v.kind = kind;
  v.nx = MF::NX;
  v.nu = MF::NU;
  int i = 0;
  for (int d : dims) v.dims[i++] = d;
  v.elems_f32 = MF::kElems;
  v.elems_f64 = MD::kElems;
  v.macs = MF::kMacs;
  v.f64_s1 = &launch_rollout<MD, double, double, double, 1>;
  v.mixed_s1 = &launch_rollout<MF, float, double, float, 1>;
  v.f32_s1 = &launch_rollout<MF, float, float, float, 1>;
  if constexpr (WITH_S2) {
    v.mixed_s2 = &launch_rollout<MF, float, double, float, 2>;
    v.f32_s2 = &launch_rollout<MF, float, float, float, 2>;
  }
  return v;
}

template <class MM64, class MM32>
void add_mma(ModelVariant& v) {
  v.bytes_mma_f64 = MM64::kBytes;
  v.bytes_mma_tf32 = MM32::kBytes;
  v.f64_mma[0] = &launch_rollout_mma<MM64, double, double, 1, true>;
  v.f64_mma[1] = &launch_rollout_mma<MM64, double, double, 2, true>;
  v.f64_mma[2] = &launch_rollout_mma<MM64, double, double, 4, false>;
  v.tf32_mma[0] = &launch_rollout_mma<MM32, float, float, 1, true>;
  v.tf32_mma[1] = &launch_rollout_mma<MM32, float, float, 2, false>;
  v.tf32_mma_wide = &launch_rollout_mma<MM32, float, float, 2, false, 256>;
  v.tf32_mma_sm = &launch_rollout_mma<MM32, float, float, 2, false, 512>;
  if constexpr (MM32::kBytes <= 100 * 1024) v.tf32_mma_mix = &launch_rollout_mma_mix<MM32, float, float>;
  if constexpr (MM32::kBytes <= 100 * 1024) v.tf32_mma_persist = &launch_rollout_mma_persist<MM32, float, float>;
  v.tf32_pipe = &launch_rollout_mma_pipe<MM32>;
}

// GP-residual kernels live in their own translation unit (variant_gp.cu)
GpLaunchers gp_launchers_rex_push();
GpLaunchers gp_launchers_rex_peg();
GpLaunchers gp_launchers_mlp_5_3_32();

template <class MM16>
void add_f16(ModelVariant& v) {
  v.bytes_mma_f16 = MM16::kBytes;
  v.f16_mma[0] = &launch_rollout_mma<MM16, float, float, 1, true>;
  v.f16_mma[1] = &launch_rollout_mma<MM16, float, float, 2, false>;
  v.f16_mma_wide = &launch_rollout_mma<MM16, float, float, 2, false, 256>;
  v.f16_mma_sm = &launch_rollout_mma<MM16, float, float, 2, false, 512>;
  if constexpr (MM16::kBytes <= 100 * 1024) v.f16_mma_mix = &launch_rollout_mma_mix<MM16, float, float>;
  if constexpr (MM16::kBytes <= 100 * 1024) v.f16_mma_persist = &launch_rollout_mma_persist<MM16, float, float>;
}

const ModelVariant* variant_rex_push();   // nx5 nu3, enc 8-16-32-5, dyn 5-32-32-5, dec 5(2)-16-32-25
const ModelVariant* variant_rex_peg();    // nx5 nu2, enc 7-16-32-5, ...
const ModelVariant* variant_mlp_5_3_32(); // 8-32-32-5
const ModelVariant* variant_mlp_5_3_64(); // 8-64-64-5
const ModelVariant* variant_mlp_5_3_128(); // 8-128-128-5 (tensor-core kernels only)
const ModelVariant* variant_pendulum();
// plain MLP (nx+nu) -> H1 -> H2 -> nx of any widths <= 256 (padded to multiples of 16), nx <= 8, nu <= 4: tcgen05 kernel (tc_rollout.cuh)
rollout_launch_fn tc_mlp_launcher(int nx, int nu);

}  // namespace tampc
