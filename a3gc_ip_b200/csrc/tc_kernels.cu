// tcgen05 / TMEM engine (placeholder until the tensor-core kernels land; reports "unsupported").
#include "common.cuh"
namespace a3gc {
bool tc_layer_supported(int, int, int, int) { return false; }
size_t tc_layer_workspace_bytes(int, int64_t, int64_t, int, int, int, int) { return 0; }
int tc_layer_forward(const LayerArgs&, void*, size_t, cudaStream_t) {
  set_error("tensor-core engine not built");
  return A3GC_ERR_UNSUPPORTED;
}
}  // namespace a3gc
