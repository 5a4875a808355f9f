// tcgen05 (3xTF32) implementation of the fused row MLP -- placeholder until the tensor-core
// tiles land: reports "unsupported" so GTB_IMPL_AUTO resolves to the FFMA tiles and an
// explicit GTB_IMPL_TCGEN05 request fails loudly.
#include "common.cuh"

namespace gtb {

bool tc_supported(int, const int32_t*) { return false; }
size_t tc_packed_bytes(int, const int32_t*) { return 0; }

int pack_tc(int, const int32_t*, const float* const*, const float* const*, void*, cudaStream_t) {
  set_error("gtb_mlp_pack: the tcgen05 layout is not available in this build");
  return GTB_ERR_UNSUPPORTED_DIM;
}

int fused_mlp_tc(const gtb_mlp_desc_t&, cudaStream_t) {
  set_error("gtb_fused_mlp_f32: the tcgen05 path is not available in this build");
  return GTB_ERR_UNSUPPORTED_DIM;
}

}  // namespace gtb
