// Pointwise (1x1) convolution as a tcgen05 / TMEM int8 GEMM -- placeholder until the
// UMMA kernel lands: reports "not taken" so vbt_detect uses the dp4a kernel in net.cu.
#include "model.cuh"

namespace vbt {
int launch_pw_umma(const vbt_model*, const OpRecord&, const int8_t*, const int8_t*, int8_t*,
                   long long, int, cudaStream_t, bool* taken) {
  *taken = false;
  return VBT_OK;
}
}  // namespace vbt
