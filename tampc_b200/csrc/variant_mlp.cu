#include "variants.cuh"
namespace tampc {
const ModelVariant* variant_mlp_5_3_32() {
  static const ModelVariant v = [] {
    ModelVariant m = make_variant<ModelMlp<float, 5, 3, 32, 32>, ModelMlp<double, 5, 3, 32, 32>, true>("mlp_nx5_nu3_h32-32", TAMPC_DYN_MLP, {32, 32});
    add_mma<ModelMlpMma<PolicyF64, 5, 3, 32, 32>, ModelMlpMma<PolicyTF32x3, 5, 3, 32, 32>>(m);
    add_f16<ModelMlpMma<PolicyF16x3, 5, 3, 32, 32>>(m);
    m.gp = gp_launchers_mlp_5_3_32();
    return m;
  }();
  return &v;
}
const ModelVariant* variant_pendulum() {
  static const ModelVariant v = make_variant<ModelPendulum<float>, ModelPendulum<double>, true>("pendulum_analytic", TAMPC_DYN_PENDULUM, {});
  return &v;
}
}  // namespace tampc
