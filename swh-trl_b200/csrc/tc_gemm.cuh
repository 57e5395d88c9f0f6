// Host entry points of the CTA-pair tcgen05 GEMM family (k7_tc_gemm.cu), for the other translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200trl {

enum TcEpilogue {
    TC_EPI_STATS = 0,      // nothing stored: per-row online-softmax statistics into `partial`
    TC_EPI_STORE = 1,      // out = bf16 [M, ldd] = D (+ bias[N]) (+ addend fp32 [M, ld_addend]); with `partial` the
                           // statistics of the rounded values as well
    TC_EPI_ACCUM = 2,      // out = fp32 [M, ldd], out += D
    TC_EPI_STORE_F32 = 4,  // out = fp32 [M, ldd], out = D
};

// D[M, N] = A[M, K] B[N, K]^T.  a_mn / b_mn = 0: operand stored [rows, k] (k contiguous); 1: stored [k, rows].
struct TcGemmParams {
    int a_mn = 0, b_mn = 0, epi = TC_EPI_STORE;
    const void* A = nullptr;
    int64_t lda = 0;
    const void* B = nullptr;
    int64_t ldb = 0;
    int64_t M = 0, N = 0, K = 0;
    void* out = nullptr;
    int64_t ldd = 0;
    const void* bias = nullptr;     // bf16 [N], TC_EPI_STORE only
    const float* addend = nullptr;  // fp32 [M, ld_addend], TC_EPI_STORE only
    int64_t ld_addend = 0;
    const int64_t* ids = nullptr;   // statistics: selected column per row
    float c = 0.f;                  // statistics: inv_T * log2(e)
    void* partial = nullptr;        // statistics workspace (tc_stats_workspace_bytes)
    int* n_groups_out = nullptr;    // statistics: number of partials per row, for tc_merge_stats
    int m_fastest = 1;              // work order without statistics: 1 = clusters running together share the B tile
                                    // (and an L2-sized super-block of A rows), 0 = they share the A tile
    void* splitk_ws = nullptr;      // TC_EPI_STORE: fp32 scratch; with it a contraction with too few output tiles is
    int64_t splitk_ws_bytes = 0;    // split along K (planes added in slice order -> deterministic)
};
int tc_gemm(const TcGemmParams& p, cudaStream_t s);
int64_t tc_stats_workspace_bytes(int64_t n_rows, int64_t n_cols);
int tc_merge_stats(const void* partial, int n_groups, int64_t n_rows, float c, float* logp, float* entropy, float* lse,
                   cudaStream_t s);

}  // namespace b200trl
