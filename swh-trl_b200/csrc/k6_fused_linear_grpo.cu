// a-13  b200trl_fused_linear_grpo : the reference's one real operator seam, `self.liger_grpo_loss(...)`
//       (trl/trainer/grpo_trainer.py:870-886 ctor, :2005-2045 call), as ONE C-ABI call.
//
// loss(hidden @ W^T + b) forward AND backward without ever holding the [B,T,V] logits: per chunk of whole sequences
//     logits_c = hidden_c W^T (+ b)          K7 CTA-pair tcgen05 GEMM, bf16 in / fp32 accumulate in TMEM / bf16 out
//     K1 resident kernel, IN PLACE           logits_c -> dlogits_c, emitting log-probs and entropies (one pass)
//     dH_c  = dlogits_c W                    K7, split-K over V with deterministic fp32 planes
//     dW   += dlogits_c^T hidden_c           K7, fp32 running sum updated by TMA reduce-add; the last chunk's epilogue
//                                            rounds bf16(dW_fp32 + D) itself (no cast pass)
// and K2 turns the log-probs into the loss value and the logged metrics.  Every contraction runs on this library's own
// tcgen05 kernel (k7_tc_gemm.cu) by default.  cuBLASLt stays available per GEMM as the A/B baseline
// (B200TRL_SEAM_GEMM / b200trl_set_seam_gemm_mask: bit 0 logits, bit 1 dH, bit 2 dW; set = ours): measured in situ
// at config 4 under the 1 kW cap (tools/seam_timeline.py) the library route takes 42.0 ms, ours 42.3 ms.  cuBLASLt
// is bound at first use with dlopen, so libb200trl.so has no link-time dependency on it (the CPU-side ABI tests load
// the library on a box without a GPU) and it is not loaded at all on the default path.
#include <cublasLt.h>
#include <dlfcn.h>

#include <cstdlib>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace b200trl {
namespace {

// ------------------------------------------------------------------ cuBLASLt, bound lazily
struct LtApi {
    void* so = nullptr;
    decltype(&cublasLtCreate) Create = nullptr;
    decltype(&cublasLtMatmul) Matmul = nullptr;
    decltype(&cublasLtMatmulDescCreate) DescCreate = nullptr;
    decltype(&cublasLtMatmulDescDestroy) DescDestroy = nullptr;
    decltype(&cublasLtMatmulDescSetAttribute) DescSet = nullptr;
    decltype(&cublasLtMatrixLayoutCreate) LayoutCreate = nullptr;
    decltype(&cublasLtMatrixLayoutDestroy) LayoutDestroy = nullptr;
    decltype(&cublasLtMatmulPreferenceCreate) PrefCreate = nullptr;
    decltype(&cublasLtMatmulPreferenceDestroy) PrefDestroy = nullptr;
    decltype(&cublasLtMatmulPreferenceSetAttribute) PrefSet = nullptr;
    decltype(&cublasLtMatmulAlgoGetHeuristic) Heuristic = nullptr;
    std::map<int, cublasLtHandle_t> handles;  // one per device
    bool ok = false;
};

std::mutex g_lt_mutex;
LtApi g_lt;

template <typename F>
bool bind(F& fn, void* so, const char* name) {
    fn = reinterpret_cast<F>(dlsym(so, name));
    return fn != nullptr;
}

LtApi* lt_api() {
    std::lock_guard<std::mutex> lock(g_lt_mutex);
    if (g_lt.ok) return &g_lt;
    if (!g_lt.so) {
        for (const char* name : {"libcublasLt.so.12", "libcublasLt.so", "/usr/local/cuda/lib64/libcublasLt.so.12"}) {
            g_lt.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (g_lt.so) break;
        }
    }
    if (!g_lt.so) {
        set_error("fused_linear_grpo: cannot load libcublasLt.so.12 (%s)", dlerror());
        return nullptr;
    }
    LtApi& a = g_lt;
    a.ok = bind(a.Create, a.so, "cublasLtCreate") && bind(a.Matmul, a.so, "cublasLtMatmul") &&
           bind(a.DescCreate, a.so, "cublasLtMatmulDescCreate") && bind(a.DescDestroy, a.so, "cublasLtMatmulDescDestroy") &&
           bind(a.DescSet, a.so, "cublasLtMatmulDescSetAttribute") &&
           bind(a.LayoutCreate, a.so, "cublasLtMatrixLayoutCreate") &&
           bind(a.LayoutDestroy, a.so, "cublasLtMatrixLayoutDestroy") &&
           bind(a.PrefCreate, a.so, "cublasLtMatmulPreferenceCreate") &&
           bind(a.PrefDestroy, a.so, "cublasLtMatmulPreferenceDestroy") &&
           bind(a.PrefSet, a.so, "cublasLtMatmulPreferenceSetAttribute") &&
           bind(a.Heuristic, a.so, "cublasLtMatmulAlgoGetHeuristic");
    if (!a.ok) {
        set_error("fused_linear_grpo: libcublasLt lacks an expected symbol");
        return nullptr;
    }
    return &a;
}

cublasLtHandle_t lt_handle(LtApi* api) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_lt_mutex);
    auto it = api->handles.find(dev);
    if (it != api->handles.end()) return it->second;
    cublasLtHandle_t h = nullptr;
    if (api->Create(&h) != CUBLAS_STATUS_SUCCESS) return nullptr;
    api->handles[dev] = h;
    return h;
}

// One column-major GEMM  D[m,n] = alpha * op(A)[m,k] op(B)[k,n] + beta * D  (+ bias[m] broadcast over columns):
// bf16 operands, fp32 accumulation, D bf16 or fp32.  Algorithms are chosen once per shape by the library's heuristic.
struct GemmKey {
    int ta, tb, dtype, bias;
    int64_t m, n, k, lda, ldb, ldd;
    bool operator<(const GemmKey& o) const {
        return std::tie(ta, tb, dtype, bias, m, n, k, lda, ldb, ldd) <
               std::tie(o.ta, o.tb, o.dtype, o.bias, o.m, o.n, o.k, o.lda, o.ldb, o.ldd);
    }
};
std::map<std::pair<int, GemmKey>, cublasLtMatmulAlgo_t> g_algos;

int lt_gemm(LtApi* api, cublasLtHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int64_t m, int64_t n, int64_t k,
            const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd, cudaDataType_t dtype, float beta,
            const void* bias, void* ws, size_t ws_bytes, cudaStream_t stream) {
    cublasLtMatmulDesc_t desc = nullptr;
    cublasLtMatrixLayout_t la = nullptr, lb = nullptr, ld = nullptr;
    cublasLtMatmulPreference_t pref = nullptr;
    int rc = B200TRL_E_LAUNCH;
    bool said = false;  // a specific message has been recorded
    const float alpha = 1.f;
    do {
        if (api->DescCreate(&desc, CUBLAS_COMPUTE_32F, CUDA_R_32F) != CUBLAS_STATUS_SUCCESS) break;
        api->DescSet(desc, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta));
        api->DescSet(desc, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb));
        if (bias) {
            const cublasLtEpilogue_t ep = CUBLASLT_EPILOGUE_BIAS;
            if (api->DescSet(desc, CUBLASLT_MATMUL_DESC_EPILOGUE, &ep, sizeof(ep)) != CUBLAS_STATUS_SUCCESS) break;
            if (api->DescSet(desc, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)) != CUBLAS_STATUS_SUCCESS) break;
        }
        const int64_t a_rows = (ta == CUBLAS_OP_N) ? m : k, a_cols = (ta == CUBLAS_OP_N) ? k : m;
        const int64_t b_rows = (tb == CUBLAS_OP_N) ? k : n, b_cols = (tb == CUBLAS_OP_N) ? n : k;
        if (api->LayoutCreate(&la, CUDA_R_16BF, a_rows, a_cols, lda) != CUBLAS_STATUS_SUCCESS) break;
        if (api->LayoutCreate(&lb, CUDA_R_16BF, b_rows, b_cols, ldb) != CUBLAS_STATUS_SUCCESS) break;
        if (api->LayoutCreate(&ld, dtype, m, n, ldd) != CUBLAS_STATUS_SUCCESS) break;
        int dev = 0;
        cudaGetDevice(&dev);
        const GemmKey key{(int)ta, (int)tb, (int)dtype, bias ? 1 : 0, m, n, k, lda, ldb, ldd};
        cublasLtMatmulAlgo_t algo;
        bool have = false;
        {
            std::lock_guard<std::mutex> lock(g_lt_mutex);
            auto it = g_algos.find({dev, key});
            if (it != g_algos.end()) {
                algo = it->second;
                have = true;
            }
        }
        if (!have) {
            if (api->PrefCreate(&pref) != CUBLAS_STATUS_SUCCESS) break;
            api->PrefSet(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof(ws_bytes));
            cublasLtMatmulHeuristicResult_t res;
            int found = 0;
            if (api->Heuristic(h, desc, la, lb, ld, ld, pref, 1, &res, &found) != CUBLAS_STATUS_SUCCESS || found == 0) {
                set_error("fused_linear_grpo: cuBLASLt has no algorithm for m=%lld n=%lld k=%lld", (long long)m,
                          (long long)n, (long long)k);
                rc = B200TRL_E_UNSUPPORTED;
                said = true;
                break;
            }
            algo = res.algo;
            std::lock_guard<std::mutex> lock(g_lt_mutex);
            g_algos[{dev, key}] = algo;
        }
        const cublasStatus_t st = api->Matmul(h, desc, &alpha, A, la, B, lb, &beta, D, ld, D, ld, &algo, ws, ws_bytes, stream);
        if (st != CUBLAS_STATUS_SUCCESS) {
            set_error("fused_linear_grpo: cublasLtMatmul failed with status %d", (int)st);
            said = true;
            break;
        }
        rc = B200TRL_OK;
    } while (false);
    if (pref) api->PrefDestroy(pref);
    if (ld) api->LayoutDestroy(ld);
    if (lb) api->LayoutDestroy(lb);
    if (la) api->LayoutDestroy(la);
    if (desc) api->DescDestroy(desc);
    if (rc != B200TRL_OK && !said) set_error("fused_linear_grpo: cuBLASLt descriptor setup failed");
    return rc;
}

// d(bias)[v] += sum over the chunk's rows of dlogits[r, v]; one thread per column, rows in order => deterministic
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ dl, int64_t rows, int64_t V,
                                                     float* __restrict__ db) {
    const int64_t v = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (v >= V) return;
    float s = 0.f;
    for (int64_t r = 0; r < rows; ++r) s += __bfloat162float(dl[r * V + v]);
    db[v] += s;
}

// bit 0 = logits GEMM, bit 1 = dH, bit 2 = dW on the CTA-pair tcgen05 kernel (k7_tc_gemm.cu); clear = cuBLASLt
int default_gemm_mask() {
    const char* v = getenv("B200TRL_SEAM_GEMM");
    return v ? (atoi(v) & 7) : 7;
}
int g_gemm_mask = default_gemm_mask();

constexpr size_t kLtWorkspace = size_t(64) << 20;
size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct Layout {
    size_t logits, row_count, total, k2, lt, end;
};
Layout layout(int64_t B, int64_t T, int64_t H, int64_t V, int64_t chunk_seqs) {
    Layout l;
    size_t off = 0;
    l.logits = off;
    off += align256(static_cast<size_t>(chunk_seqs) * T * V * 2);
    l.row_count = off;
    off += align256(static_cast<size_t>(B) * 4);
    l.total = off;
    off += 256;
    l.k2 = off;
    off += align256(static_cast<size_t>(b200trl_grpo_loss_workspace_bytes(B)));
    l.lt = off;  // cuBLASLt scratch, or the fp32 split-K planes of the tcgen05 dH GEMM (never both at once)
    off += std::max<size_t>(kLtWorkspace, align256(static_cast<size_t>(
                                              b200trl_tc_gemm_workspace_bytes(chunk_seqs * T, H, V, B200TRL_TC_OUT_BF16))));
    l.end = off;
    return l;
}

}  // namespace
}  // namespace b200trl

using namespace b200trl;

extern "C" int b200trl_set_seam_gemm_mask(int mask) {
    const int prev = g_gemm_mask;
    if (mask >= 0) g_gemm_mask = mask & 7;
    return prev;
}

extern "C" int64_t b200trl_fused_linear_grpo_workspace_bytes(int64_t B, int64_t T, int64_t H, int64_t V,
                                                             int64_t chunk_seqs) {
    if (B <= 0 || T <= 0 || H <= 0 || V <= 0 || chunk_seqs <= 0) return 0;
    return static_cast<int64_t>(layout(B, T, H, V, std::min(chunk_seqs, B)).end);
}

// fp32 -> bf16, 8 elements per thread: only needed when the LIBRARY produced the last chunk of dW (our own kernel
// rounds inside its epilogue)
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                            int64_t n8) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n8) return;
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    __nv_bfloat162 r[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w),
                           __floats2bfloat162_rn(b.x, b.y), __floats2bfloat162_rn(b.z, b.w)};
    reinterpret_cast<uint4*>(dst)[i] = *reinterpret_cast<uint4*>(r);
}

static int seam_impl(const void* hidden, const void* weight, const void* bias, int64_t B, int64_t T, int64_t H, int64_t V,
                     const int64_t* ids, const int32_t* mask, const float* advantages, const float* old_logp,
                     const float* ref_logp, const b200trl_grpo_cfg* cfg, float inv_temperature, int64_t chunk_seqs,
                     void* workspace, float* logp, float* entropy, float* loss, float* metrics, void* dhidden,
                     float* dweight, void* dweight_bf16, float* dbias, const int64_t* seq_rows, b200trl_stream_t stream_);

extern "C" int b200trl_fused_linear_grpo(const void* hidden, const void* weight, const void* bias, int64_t B, int64_t T,
                                         int64_t H, int64_t V, const int64_t* ids, const int32_t* mask,
                                         const float* advantages, const float* old_logp, const float* ref_logp,
                                         const b200trl_grpo_cfg* cfg, float inv_temperature, int64_t chunk_seqs,
                                         void* workspace, float* logp, float* entropy, float* loss, float* metrics,
                                         void* dhidden, float* dweight, void* dweight_bf16, float* dbias,
                                         b200trl_stream_t stream_) {
    return seam_impl(hidden, weight, bias, B, T, H, V, ids, mask, advantages, old_logp, ref_logp, cfg, inv_temperature,
                     chunk_seqs, workspace, logp, entropy, loss, metrics, dhidden, dweight, dweight_bf16, dbias, nullptr,
                     stream_);
}

// The same operator with the padding trimmed: seq_rows_host[b] (HOST memory, B entries, 0 <= value <= T) is the number of
// leading rows of sequence b that can carry a non-zero mask (index of its last unmasked token + 1).  The rows behind it
// take part in none of the three contractions: every sequence is its own chunk of seq_rows[b] rows, so the GEMMs do
// sum(seq_rows) / (B T) of the dense work.  Results are those of the dense call: loss, metrics, dW, db identical up to
// the order of the fp32 sums, dH rows behind seq_rows[b] are zeros (as their dlogits are), logp / entropy there are 0.
extern "C" int b200trl_fused_linear_grpo_trimmed(const void* hidden, const void* weight, const void* bias, int64_t B,
                                                 int64_t T, int64_t H, int64_t V, const int64_t* ids, const int32_t* mask,
                                                 const float* advantages, const float* old_logp, const float* ref_logp,
                                                 const b200trl_grpo_cfg* cfg, float inv_temperature,
                                                 const int64_t* seq_rows_host, void* workspace, float* logp, float* entropy,
                                                 float* loss, float* metrics, void* dhidden, float* dweight,
                                                 void* dweight_bf16, float* dbias, b200trl_stream_t stream_) {
    B200TRL_REQUIRE(seq_rows_host != nullptr, B200TRL_E_INVALID, "fused_linear_grpo_trimmed: null seq_rows");
    for (int64_t b = 0; b < B; ++b)
        B200TRL_REQUIRE(seq_rows_host[b] >= 0 && seq_rows_host[b] <= T, B200TRL_E_INVALID,
                        "fused_linear_grpo_trimmed: seq_rows[%lld] = %lld outside [0, T]", (long long)b,
                        (long long)seq_rows_host[b]);
    return seam_impl(hidden, weight, bias, B, T, H, V, ids, mask, advantages, old_logp, ref_logp, cfg, inv_temperature, 1,
                     workspace, logp, entropy, loss, metrics, dhidden, dweight, dweight_bf16, dbias, seq_rows_host, stream_);
}

static int seam_impl(const void* hidden, const void* weight, const void* bias, int64_t B, int64_t T, int64_t H, int64_t V,
                     const int64_t* ids, const int32_t* mask, const float* advantages, const float* old_logp,
                     const float* ref_logp, const b200trl_grpo_cfg* cfg, float inv_temperature, int64_t chunk_seqs,
                     void* workspace, float* logp, float* entropy, float* loss, float* metrics, void* dhidden,
                     float* dweight, void* dweight_bf16, float* dbias, const int64_t* seq_rows, b200trl_stream_t stream_) {
    B200TRL_REQUIRE(hidden && weight && ids && mask && advantages && cfg && workspace && logp && entropy && loss && metrics,
                    B200TRL_E_INVALID, "fused_linear_grpo: null pointer");
    B200TRL_REQUIRE(B > 0 && T > 0 && H > 0 && V > 0 && chunk_seqs > 0, B200TRL_E_INVALID, "fused_linear_grpo: bad shape");
    B200TRL_REQUIRE(H % 8 == 0 && V % 8 == 0, B200TRL_E_UNSUPPORTED,
                    "fused_linear_grpo: hidden size and vocabulary must be multiples of 8 (16-byte rows)");
    B200TRL_REQUIRE(!dbias || bias, B200TRL_E_INVALID, "fused_linear_grpo: dbias without bias");
    B200TRL_REQUIRE(cfg->is_level == B200TRL_IS_TOKEN || old_logp == nullptr, B200TRL_E_UNSUPPORTED,
                    "fused_linear_grpo: sequence-level importance sampling with old_logp needs the two-phase path");
    B200TRL_REQUIRE(cfg->beta == 0.f || ref_logp, B200TRL_E_INVALID, "fused_linear_grpo: beta != 0 needs ref_logp");
    chunk_seqs = std::min(chunk_seqs, B);
    int64_t n_chunks = (B + chunk_seqs - 1) / chunk_seqs;
    if (seq_rows) {  // one chunk per sequence that has rows at all
        n_chunks = 0;
        for (int64_t b = 0; b < B; ++b) n_chunks += seq_rows[b] > 0 ? 1 : 0;
    }
    B200TRL_REQUIRE(!dweight_bf16 || dweight || n_chunks == 1, B200TRL_E_INVALID,
                    "fused_linear_grpo: dweight_bf16 over several chunks needs the fp32 dweight accumulator as well");
    const bool want_dw = dweight || dweight_bf16;
    cudaStream_t stream = as_stream(stream_);
    const int mask_tc = g_gemm_mask;
    const bool need_lt = !(mask_tc & 1) || (dhidden && !(mask_tc & 2)) || (want_dw && !(mask_tc & 4));
    LtApi* api = need_lt ? lt_api() : nullptr;
    if (need_lt && !api) return B200TRL_E_UNSUPPORTED;
    cublasLtHandle_t h = api ? lt_handle(api) : nullptr;
    B200TRL_REQUIRE(!need_lt || h != nullptr, B200TRL_E_LAUNCH, "fused_linear_grpo: cublasLtCreate failed");

    const Layout l = layout(B, T, H, V, chunk_seqs);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    __nv_bfloat16* logits = reinterpret_cast<__nv_bfloat16*>(ws + l.logits);
    float* row_count = reinterpret_cast<float*>(ws + l.row_count);
    float* total = reinterpret_cast<float*>(ws + l.total);
    void* k2_ws = ws + l.k2;
    void* lt_ws = ws + l.lt;
    const bool want_grad = dhidden || want_dw || dbias;

    int rc = b200trl_mask_stats(mask, B, T, row_count, total, stream_);
    if (rc) return rc;
    if (cudaMemsetAsync(k2_ws, 0, static_cast<size_t>(b200trl_grpo_loss_workspace_bytes(B)), stream) != cudaSuccess)
        return check_launch("fused_linear_grpo memset");
    if (dbias && cudaMemsetAsync(dbias, 0, static_cast<size_t>(V) * 4, stream) != cudaSuccess)
        return check_launch("fused_linear_grpo memset");
    // dW is not zeroed: the first chunk's contraction writes it (beta = 0), the following ones accumulate

    const __nv_bfloat16* hid = static_cast<const __nv_bfloat16*>(hidden);
    __nv_bfloat16* dh = static_cast<__nv_bfloat16*>(dhidden);
    if (seq_rows && n_chunks == 0) {  // nothing unmasked anywhere: every gradient is zero
        if (dweight && cudaMemsetAsync(dweight, 0, static_cast<size_t>(V) * H * 4, stream) != cudaSuccess)
            return check_launch("fused_linear_grpo memset");
        if (dweight_bf16 && cudaMemsetAsync(dweight_bf16, 0, static_cast<size_t>(V) * H * 2, stream) != cudaSuccess)
            return check_launch("fused_linear_grpo memset");
    }
    int64_t chunk = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk_seqs) {
        const int64_t nb = std::min(chunk_seqs, B - b0), r0 = b0 * T;
        int64_t rows = nb * T;
        if (seq_rows) {  // trimmed: this sequence's leading rows only; what lies behind them is defined here
            rows = seq_rows[b0];
            const size_t tail = static_cast<size_t>(T - rows);
            if (tail) {
                if (cudaMemsetAsync(logp + r0 + rows, 0, tail * 4, stream) != cudaSuccess ||
                    cudaMemsetAsync(entropy + r0 + rows, 0, tail * 4, stream) != cudaSuccess ||
                    (dh && cudaMemsetAsync(dh + (r0 + rows) * H, 0, tail * H * 2, stream) != cudaSuccess))
                    return check_launch("fused_linear_grpo memset");
            }
            if (rows == 0) continue;
        }
        const bool first = chunk == 0, last = chunk == n_chunks - 1;
        ++chunk;
        // row-major logits[rows, V] = hidden_c[rows, H] W[V, H]^T  <=>  column-major D[V, rows] = W^T(T) x hidden_c(N)
        if (mask_tc & 1) {  // hidden_c and W both K-major; all row blocks of one W tile run at the same time
            TcGemmParams p;
            p.A = hid + r0 * H, p.lda = H, p.B = weight, p.ldb = H, p.M = rows, p.N = V, p.K = H;
            p.out = logits, p.ldd = V, p.bias = bias, p.m_fastest = 1;
            rc = tc_gemm(p, stream);
        } else {
            rc = lt_gemm(api, h, CUBLAS_OP_T, CUBLAS_OP_N, V, rows, H, weight, H, hid + r0 * H, H, logits, V, CUDA_R_16BF,
                         0.f, bias, lt_ws, kLtWorkspace, stream);
        }
        if (rc) return rc;
        // the loss normalises over the WHOLE batch: grpo / dr_grpo divide by B (grpo_trainer.py:2131, 2135), bnpo by
        // the batch's token total, which the kernel reads from `total`
        b200trl_grpo_cfg c = *cfg;
        c.grad_scale = cfg->grad_scale * (cfg->loss_type == B200TRL_LOSS_BNPO ? 1.f
                                                                              : static_cast<float>(nb) / static_cast<float>(B));
        rc = b200trl_grpo_fused_fwd_bwd(logits, B200TRL_BF16, nb, seq_rows ? rows : T, V, V, 0, ids + r0, mask + r0,
                                        advantages + b0,
                                        old_logp ? old_logp + r0 : nullptr, ref_logp ? ref_logp + r0 : nullptr, &c,
                                        inv_temperature, row_count + b0, total, logp + r0, entropy + r0, nullptr,
                                        want_grad ? logits : nullptr, V, 0, stream_);  // in place: dlogits overwrite logits
        if (rc) return rc;
        if (dh) {  // dH_c[rows, H] = dl[rows, V] W[V, H]  <=>  D[H, rows] = W(N)[H, V] x dl(N)[V, rows]
            if (mask_tc & 2) {  // A = dl (K = V contiguous), B[n = h, k = v] = W[v, h] is MN-major; few tiles -> split-K
                TcGemmParams p;
                p.b_mn = 1;
                p.A = logits, p.lda = V, p.B = weight, p.ldb = H, p.M = rows, p.N = H, p.K = V;
                p.out = dh + r0 * H, p.ldd = H, p.m_fastest = 1;
                p.splitk_ws = lt_ws, p.splitk_ws_bytes = static_cast<int64_t>(l.end - l.lt);
                rc = tc_gemm(p, stream);
            } else {
                rc = lt_gemm(api, h, CUBLAS_OP_N, CUBLAS_OP_N, H, rows, V, weight, H, logits, V, dh + r0 * H, H, CUDA_R_16BF,
                             0.f, nullptr, lt_ws, kLtWorkspace, stream);
            }
            if (rc) return rc;
        }
        if (want_dw) {  // dW[V, H] (+)= dl^T hidden_c  <=>  D[H, V] (+)= hidden_c(N)[H, rows] x dl(T)[rows, V]
            if (mask_tc & 4) {
                // A[m = v, k = r] = dl[r, v] and B[n = h, k = r] = hidden[r, h]: both MN-major; the dl tile is the
                // operand that does not fit L2, so clusters running together share it.  First chunk: plain fp32 store;
                // middle chunks: fp32 accumulate; last chunk with a bf16 destination: bf16(dW_fp32 + D) in the epilogue
                TcGemmParams p;
                p.a_mn = p.b_mn = 1;
                p.A = logits, p.lda = V, p.B = hid + r0 * H, p.ldb = H, p.M = V, p.N = H, p.K = rows;
                p.m_fastest = 0;
                if (last && dweight_bf16) {
                    p.epi = TC_EPI_STORE, p.out = dweight_bf16, p.ldd = H;
                    if (!first) p.addend = dweight, p.ld_addend = H;
                } else {
                    p.epi = first ? TC_EPI_STORE_F32 : TC_EPI_ACCUM, p.out = dweight, p.ldd = H;
                }
                rc = tc_gemm(p, stream);
            } else {
                rc = lt_gemm(api, h, CUBLAS_OP_N, CUBLAS_OP_T, H, V, rows, hid + r0 * H, H, logits, V, dweight, H, CUDA_R_32F,
                             first ? 0.f : 1.f, nullptr, lt_ws, kLtWorkspace, stream);
                if (!rc && last && dweight_bf16) {
                    const int64_t n8 = V * H / 8;
                    cast_f32_bf16_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(
                        dweight, static_cast<__nv_bfloat16*>(dweight_bf16), n8);
                    rc = check_launch("cast_f32_bf16_kernel");
                }
            }
            if (rc) return rc;
        }
        if (dbias) {
            colsum_kernel<<<static_cast<unsigned>((V + 255) / 256), 256, 0, stream>>>(logits, rows, V, dbias);
            rc = check_launch("colsum_kernel");
            if (rc) return rc;
        }
    }
    b200trl_grpo_cfg c = *cfg;
    c.grad_scale = 1.f;
    return b200trl_grpo_loss(logp, old_logp, ref_logp, advantages, mask, nullptr, entropy, B, T, &c, row_count, total, k2_ws,
                             loss, metrics, nullptr, stream_);
}
