// variants.cuh — the compiled dynamics shapes.  Each lives in its own translation unit (parallel nvcc).
#pragma once
#include <initializer_list>
#include "rollout_kernel.cuh"

namespace tampc {

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
  v.f64_s1 = &launch_rollout<MD, double, double, 1>;
  v.mixed_s1 = &launch_rollout<MF, float, double, 1>;
  v.f32_s1 = &launch_rollout<MF, float, float, 1>;
  if constexpr (WITH_S2) {
    v.mixed_s2 = &launch_rollout<MF, float, double, 2>;
    v.f32_s2 = &launch_rollout<MF, float, float, 2>;
  } else {
    v.mixed_s2 = nullptr;
    v.f32_s2 = nullptr;
  }
  return v;
}

const ModelVariant* variant_rex_push();   // nx5 nu3, enc 8-16-32-5, dyn 5-32-32-5, dec 5(2)-16-32-25
const ModelVariant* variant_rex_peg();    // nx5 nu2, enc 7-16-32-5, ...
const ModelVariant* variant_mlp_5_3_32(); // 8-32-32-5
const ModelVariant* variant_mlp_5_3_64(); // 8-64-64-5
const ModelVariant* variant_pendulum();

}  // namespace tampc
