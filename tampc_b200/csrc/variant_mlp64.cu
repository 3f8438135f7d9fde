#include "variants.cuh"
namespace tampc {
const ModelVariant* variant_mlp_5_3_64() {
  static const ModelVariant v = [] {
    ModelVariant m = make_variant<ModelMlp<float, 5, 3, 64, 64>, ModelMlp<double, 5, 3, 64, 64>, false>("mlp_nx5_nu3_h64-64", TAMPC_DYN_MLP, {64, 64});
    add_mma<ModelMlpMma<PolicyF64, 5, 3, 64, 64>, ModelMlpMma<PolicyTF32x3, 5, 3, 64, 64>>(m);
    add_f16<ModelMlpMma<PolicyF16x3, 5, 3, 64, 64>>(m);
    return m;
  }();
  return &v;
}
}  // namespace tampc
