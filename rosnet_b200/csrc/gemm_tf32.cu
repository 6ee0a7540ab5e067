// float32 / complex64 grouped block GEMM (error-compensated 3xTF32 on tcgen05, fp32 TMEM
// accumulators).  Placeholder translation unit until the kernel lands: fails loudly.
#include "common.h"

namespace rnb {
int contract_tf32_workspace(const rnb_contract_desc *d, size_t *bytes) {
  (void)d;
  *bytes = 0;
  return RNB_OK;
}
int contract_tf32(const rnb_contract_desc *d, cudaStream_t stream) {
  (void)d;
  (void)stream;
  set_error("float32/complex64 contraction (3xTF32 tcgen05 kernel) is not built into this library yet");
  return RNB_ERR_UNSUPPORTED;
}
}  // namespace rnb
