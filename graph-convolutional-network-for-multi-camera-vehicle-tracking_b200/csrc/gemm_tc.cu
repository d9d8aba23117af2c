// tcgen05 3xTF32 GEMM (placeholder until the TMA/TMEM kernel lands): reports "unsupported" so callers use the SIMT kernel.
#include "common.cuh"
#include "kernels.h"

namespace mpn {
size_t gemm_tc_workspace_bytes(int, int, int) { return 0; }
bool gemm_tc_supported(int, int, int) { return false; }
int gemm_nt_tc(const float*, const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return MPN_ERR_INVALID;
}
}  // namespace mpn
