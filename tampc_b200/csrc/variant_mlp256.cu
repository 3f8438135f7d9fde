#include "variants.cuh"
#include "mma_stream.cuh"
namespace tampc {
// 8 -> 256 -> 256 -> 5: the H x H layer (512 KB of split weights) is streamed from L2 (mma_stream.cuh); 3xTF32 only
const ModelVariant* variant_mlp_5_3_256() {
  static const ModelVariant v = [] {
    using M = ModelMlpStream<5, 3, 256>;
    ModelVariant m{};
    m.name = "mlp_nx5_nu3_h256-256";
    m.kind = TAMPC_DYN_MLP;
    m.nx = 5;
    m.nu = 3;
    m.dims[0] = 256;
    m.dims[1] = 256;
    m.macs = M::kMacs;
    m.bytes_mma_tf32 = M::kBytes;
    m.tf32_mma[0] = &launch_rollout_mma_stream<M>;
    return m;
  }();
  return &v;
}
}  // namespace tampc
