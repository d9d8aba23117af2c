// Internal (non-ABI) declarations shared between the translation units of libmpn_b200.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mpn {

// edge features (inference.py:453-456): eps of F.pairwise_distance / F.cosine_similarity, and the cancellation threshold below
// which a pair is recomputed directly from its two rows (edge_features.cu)
constexpr float PAIRWISE_EPS = 1e-6f;
constexpr float COSINE_EPS = 1e-8f;
constexpr float REFINE_FRACTION = 0.25f;

// EXPERIMENTAL (mpn_set_fused_distance): the distance epilogue inside the Gram GEMM.  For graphs whose rows are "all columns
// but one contiguous gap" (dense cross-camera graphs: the gap is the node's own camera) the edge id of the ordered pair
// (i, j) is closed-form, e = rowptr[i] + j - (j past the gap ? gap length : 0), so the epilogue warps turn the accumulator
// straight into edge_attr rows: no Gram matrix in HBM, no gather pass.  Whether a graph has that shape is decided on the
// device (gap_table_kernel verifies every row and raises *not_one_gap otherwise): the kernel then falls back to storing the
// Gram block, and the gather kernel that follows — skipped when the flag is clear — does the work as before.
struct EfEpilogue {
  const float4* st;        // [n_cols] per-node statistics of the centred rows (center_rows_kernel), global node ids
  const int* rowptr;       // [n_rows+1] of the graph (local rows)
  const int2* gap;         // [n_rows] (first column of the gap, length of the gap) of each local row, global column ids
  const int* not_one_gap;  // device flag: 0 = every row has the one-gap shape (fused epilogue), else store the Gram block
  float2* edge_attr;       // [E]
  int* refine_list;
  int* refine_count;
  int row_local0;          // local row / global node id of row 0 of the A block
  int row_global0;
  int D;
};
struct EfNone {};

// gemm_simt.cu
int gemm_nt_simt(const float* A, const float* B, const float* bias, const float* a_scale, const float* a_shift,
                 float* C, int M, int N, int K, cudaStream_t st, const int* row_gid = nullptr);

// gemm_tc.cu  (tcgen05 3xTF32; operands are pre-split hi/lo planes)
size_t gemm_tc_workspace_bytes(int M, int N, int K);
// a_scale/a_shift (optional, [K]): A is read as relu(A*scale + shift) while it is split (fused BatchNorm+ReLU)
int gemm_nt_tc(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
               void* workspace, size_t workspace_bytes, cudaStream_t st, const float* a_scale = nullptr,
               const float* a_shift = nullptr, const float* b_hi_cached = nullptr, const float* b_lo_cached = nullptr,
               const int* row_gid = nullptr);   // row_gid: a_scale/a_shift are [n_graphs][K] tables indexed by the row's graph
int split_tf32(const float* x, long long n, float* hi, float* lo, cudaStream_t st);
// Gram block of the row block [a_row0, a_row0+M) of X [N,K] against all of X: 3xFP16 planes scaled by max|X| (amax_dev: float
// bits on the device) when K % 8 == 0, else / when amax_dev is NULL the TF32 path of gemm_nt_tc
int gram_nt_tc(const float* X, int a_row0, float* C, int M, int N, int K, const float* amax_dev, void* workspace, size_t workspace_bytes,
               cudaStream_t st, const EfEpilogue* ef = nullptr);   // ef: fused distance epilogue (fp16 planes only; C = its fallback)
bool gram_ef_supported(int M, int N, int K, const float* amax_dev);
bool gemm_tc_supported(int M, int N, int K);
// 3xFP16 variant for the node encoder: A is split here into fp16 planes (fused BatchNorm+ReLU; plane scale from a_amax_host, or
// measured on the device when a_amax_host <= 0), B planes (b_hi16/b_lo16, scaled by b_scale) are cached by the caller
int gemm_nt_tc_f16(const float* A, const float* bias, float* C, int M, int N, int K, void* workspace, size_t workspace_bytes,
                   cudaStream_t st, const float* a_scale, const float* a_shift, const int* row_gid, float a_amax_host,
                   const void* b_hi16, const void* b_lo16, float b_scale);
int split_f16_host_scale(const float* x, long long n, float amax, void* hi, void* lo, float* scale_out, cudaStream_t st);
int gram_blockdiag_tc(const float* X, int N, int K, const int* graph_nptr, const long long* g_off, int n_graphs, int max_ng,
                      float* Gbuf, void* ws, size_t ws_bytes, cudaStream_t st, const float* amax_dev = nullptr);

}  // namespace mpn
