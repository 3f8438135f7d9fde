#include "variants.cuh"
namespace tampc {
GpLaunchers gp_launchers_rex_push() {
  return make_gp_launchers<ModelRex<float, 5, 3, 16, 32, 5, 32, 32, 5, 16, 32>, ModelRex<double, 5, 3, 16, 32, 5, 32, 32, 5, 16, 32>>();
}
GpLaunchers gp_launchers_rex_peg() {
  return make_gp_launchers<ModelRex<float, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>, ModelRex<double, 5, 2, 16, 32, 5, 32, 32, 5, 16, 32>>();
}
GpLaunchers gp_launchers_mlp_5_3_32() { return make_gp_launchers<ModelMlp<float, 5, 3, 32, 32>, ModelMlp<double, 5, 3, 32, 32>>(); }
}  // namespace tampc
