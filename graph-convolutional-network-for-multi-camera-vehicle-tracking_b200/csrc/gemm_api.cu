// C ABI of the GEMM building block (tests + roofline bench).
#include "common.cuh"
#include "kernels.h"

using namespace mpn;

extern "C" {

size_t mpn_gemm_nt_workspace_bytes(int32_t M, int32_t N, int32_t K, int impl) {
  return impl == 1 ? gemm_tc_workspace_bytes(M, N, K) + 256 : 256;
}

int mpn_gemm_nt(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K, int impl,
                void* ws, size_t ws_bytes, void* stream) {
  MPN_REQUIRE(A && B && C, "gemm: NULL argument");
  if (impl == 0) return gemm_nt_simt(A, B, bias, nullptr, nullptr, C, M, N, K, (cudaStream_t)stream);
  MPN_REQUIRE(impl == 1, "gemm: impl must be 0 (simt) or 1 (tcgen05)");
  MPN_REQUIRE(gemm_tc_supported(M, N, K), "gemm: shape %d x %d x %d not supported by the tcgen05 kernel", M, N, K);
  return gemm_nt_tc(A, B, bias, C, M, N, K, ws, ws_bytes, (cudaStream_t)stream);
}

// Gram block of rows [row0, row0+M) of X [N,K] against all of X, the way the edge features take it: 3xFP16 planes scaled from
// *amax_dev = max |X| (K % 8 == 0), symmetric tiles when the block is all of X.  amax_dev NULL -> the 3xTF32 path.
int mpn_gram_nt(const float* X, int32_t row0, float* C, int32_t M, int32_t N, int32_t K, const float* amax_dev, void* ws, size_t ws_bytes,
                void* stream) {
  MPN_REQUIRE(X && C && row0 >= 0 && row0 + M <= N, "gram: bad argument");
  MPN_REQUIRE(gemm_tc_supported(M, N, K), "gram: shape %d x %d x %d not supported by the tcgen05 kernel", M, N, K);
  return gram_nt_tc(X, row0, C, M, N, K, amax_dev, ws, ws_bytes, (cudaStream_t)stream);
}

int mpn_split_f16(const float* x, int64_t n, float amax, void* hi, void* lo, float* scale_out_host, void* stream) {
  return split_f16_host_scale(x, (long long)n, amax, hi, lo, scale_out_host, (cudaStream_t)stream);
}

int mpn_split_tf32(const float* x, int64_t n, float* hi, float* lo, void* stream) {
  return split_tf32(x, (long long)n, hi, lo, (cudaStream_t)stream);
}

}  // extern "C"
