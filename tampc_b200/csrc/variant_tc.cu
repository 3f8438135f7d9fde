#include "variants.cuh"
#include "tc_rollout.cuh"
namespace tampc {
// state / action widths the tcgen05 kernel is instantiated for (the hidden widths are run-time values)
rollout_launch_fn tc_mlp_launcher(int nx, int nu) {
#define TC_CASE(X, U) if (nx == X && nu == U) return &launch_tc_mlp<X, U>;
  TC_CASE(5, 3) TC_CASE(5, 2) TC_CASE(2, 1) TC_CASE(3, 1) TC_CASE(4, 2) TC_CASE(4, 1) TC_CASE(6, 2) TC_CASE(6, 3) TC_CASE(7, 3) TC_CASE(8, 4)
#undef TC_CASE
  return nullptr;
}
}  // namespace tampc
