#include "variants.cuh"
namespace tampc {
const ModelVariant* variant_rex_peg() {
  static const ModelVariant v = [] {
    ModelVariant m = make_variant<ModelRex<float, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>, ModelRex<double, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>, true>(
        "rex_nx5_nu2_e16-32-5_d32-32-5_c16-32", TAMPC_DYN_REX, {16, 32, 5, 32, 32, 5, 16, 32});
    add_mma<ModelRexMma<PolicyF64, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>, ModelRexMma<PolicyTF32x3, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>>(m);
    m.tf32_split = &launch_rollout_mma_split<ModelRexMma<PolicyTF32x3, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>>;
    add_f16<ModelRexMma<PolicyF16x3, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>>(m);
    m.f16_split = &launch_rollout_mma_split<ModelRexMma<PolicyF16x3, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>>;
    m.tf32_split3 = &launch_rollout_mma_split<ModelRexMma<PolicyTF32x3, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>, 3>;
    m.f16_split3 = &launch_rollout_mma_split<ModelRexMma<PolicyF16x3, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>, 3>;
    m.gp = gp_launchers_rex_peg();
    return m;
  }();
  return &v;
}
}  // namespace tampc
