#include "generic_rollout.cuh"
namespace tampc {
// every state / action width the ABI allows (TAMPC_MAX_NX x TAMPC_MAX_NU); the layer widths are run-time values
template <int NX, int NU>
static rollout_launch_fn pick_types(int types) {
  switch (types) {
    case 0: return &launch_rollout_generic<double, double, double, NX, NU>;
    case 1: return &launch_rollout_generic<float, double, float, NX, NU>;
    default: return &launch_rollout_generic<float, float, float, NX, NU>;
  }
}
template <int NX>
static rollout_launch_fn pick_nu(int nu, int types) {
  switch (nu) {
    case 1: return pick_types<NX, 1>(types);
    case 2: return pick_types<NX, 2>(types);
    case 3: return pick_types<NX, 3>(types);
    case 4: return pick_types<NX, 4>(types);
    default: return nullptr;
  }
}
rollout_launch_fn generic_launcher(int nx, int nu, int types) {
  static_assert(TAMPC_MAX_NX == 8 && TAMPC_MAX_NU == 4, "instantiation list below");
  switch (nx) {
    case 1: return pick_nu<1>(nu, types);
    case 2: return pick_nu<2>(nu, types);
    case 3: return pick_nu<3>(nu, types);
    case 4: return pick_nu<4>(nu, types);
    case 5: return pick_nu<5>(nu, types);
    case 6: return pick_nu<6>(nu, types);
    case 7: return pick_nu<7>(nu, types);
    case 8: return pick_nu<8>(nu, types);
    default: return nullptr;
  }
}
}  // namespace tampc
