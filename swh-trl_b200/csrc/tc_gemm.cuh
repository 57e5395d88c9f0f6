// Host entry points of the CTA-pair tcgen05 GEMM family (k7_tc_gemm.cu), for the other translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200trl {

enum TcEpilogue { TC_EPI_STATS = 0, TC_EPI_STORE = 1, TC_EPI_ACCUM = 2 };

// D[M, N] = A[M, K] B[N, K]^T.  a_mn / b_mn = 0: operand stored [rows, k] (k contiguous); 1: stored [k, rows].
//   TC_EPI_STATS: `partial` (tc_stats_workspace_bytes) receives per-group row statistics, nothing else is written;
//   TC_EPI_STORE: out = bf16 [M, ldd] (+ bias[N]); with `partial` the statistics of the rounded values as well;
//   TC_EPI_ACCUM: out = fp32 [M, ldd], accumulated into.
// `splitk_ws` (nullable): fp32 scratch for TC_EPI_STORE; when the output has too few tiles to fill the machine, K is
// split into slices whose partial products go to planes of the scratch and are added in slice order (deterministic).
// `m_fastest`: work order when no statistics are wanted (1: consecutive clusters share the B tile, 0: the A tile).
int tc_gemm(int a_mn, int b_mn, int epi, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N,
            int64_t K, void* out, int64_t ldd, const void* bias, const int64_t* ids, float c, void* partial,
            int* n_groups_out, int m_fastest, void* splitk_ws, int64_t splitk_ws_bytes, cudaStream_t s);
int64_t tc_stats_workspace_bytes(int64_t n_rows, int64_t n_cols);
int tc_merge_stats(const void* partial, int n_groups, int64_t n_rows, float c, float* logp, float* entropy, float* lse,
                   cudaStream_t s);

}  // namespace b200trl
