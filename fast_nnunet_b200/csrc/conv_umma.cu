// tcgen05 implicit-GEMM back end — placeholder until the kernel lands (next commit).
#include "common.cuh"
#include "ops.cuh"

namespace fnnu {
bool umma_supported(const ConvArgs&) { return false; }
size_t umma_packed_weight_bytes(int, int, int, int) { return 0; }
int launch_pack_weights_umma(const float*, void*, int, int, const int*, int, cudaStream_t) { return FNNU_OK; }
int launch_conv_umma(const ConvArgs&, cudaStream_t) {
  set_error("tcgen05 back end not built");
  return FNNU_E_UNSUPPORTED;
}
}  // namespace fnnu
