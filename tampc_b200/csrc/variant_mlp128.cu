#include "variants.cuh"
namespace tampc {
// 8 -> 128 -> 128 -> 5: tensor-core kernels only (a thread cannot hold 128-wide activations for the FMA path)
const ModelVariant* variant_mlp_5_3_128() {
  static const ModelVariant v = [] {
    ModelVariant m{};
    m.name = "mlp_nx5_nu3_h128-128";
    m.kind = TAMPC_DYN_MLP;
    m.nx = 5;
    m.nu = 3;
    m.dims[0] = 128;
    m.dims[1] = 128;
    m.macs = 8 * 128 + 128 * 128 + 128 * 5;
    add_mma<ModelMlpMma<PolicyF64, 5, 3, 128, 128>, ModelMlpMma<PolicyTF32x3, 5, 3, 128, 128>>(m);
    add_f16<ModelMlpMma<PolicyF16x3, 5, 3, 128, 128>>(m);
    m.f64_mma[2] = nullptr;  // 32 samples per warp in float64 spills: 8 / 16 samples per warp only
    return m;
  }();
  return &v;
}
}  // namespace tampc
