// C-ABI entry points of K1 (log-prob / entropy / dlogits) — argument checks and path selection.
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "k1_args.cuh"

using namespace b200trl;

namespace {

std::atomic<int> g_k1_path{B200TRL_K1_AUTO};
std::atomic<int> g_skip_masked{0};
std::atomic<unsigned long long*> g_trace{nullptr};
std::atomic<int> g_trace_row0{0};

// which implementation a call goes to (the in-kernel loss sums of b200trl_grpo_fused_step exist on the resident one)
bool takes_resident(const K1Args& a, int dtype) {
    const int path = g_k1_path.load();
    if (path == B200TRL_K1_RESIDENT) return true;
    return path == B200TRL_K1_AUTO && k1_resident_preferred(a, dtype);
}

int dispatch(K1Args a, int dtype, cudaStream_t stream) {
    a.elem_f16 = (dtype == B200TRL_F16) ? 1 : 0;
    a.trace = g_trace.load();
    a.trace_row0 = g_trace_row0.load();
    const int path = g_k1_path.load();
    const bool resident_ok = k1_resident_supported(a, dtype);
    if (path == B200TRL_K1_RESIDENT) {
        B200TRL_REQUIRE(resident_ok, B200TRL_E_UNSUPPORTED,
                        "k1: resident path needs bf16 logits, rows of >= 16384 elements and a dlogits layout that mirrors "
                        "the logits' 16-byte alignment");
        return launch_k1_resident(a, stream);
    }
    if (path == B200TRL_K1_AUTO && k1_resident_preferred(a, dtype)) return launch_k1_resident(a, stream);
    a.step_ws = nullptr;  // the row kernel has no in-kernel loss sums
    return launch_k1_row(a, dtype, stream);
}

int fill_common(K1Args& a, const void* logits, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                int64_t rows_per_batch, int64_t batch_stride, const int64_t* ids, float inv_temperature,
                const char* who) {
    B200TRL_REQUIRE(logits && ids, B200TRL_E_INVALID, "%s: null pointer", who);
    B200TRL_REQUIRE(dtype_size(dtype) != 0, B200TRL_E_UNSUPPORTED, "%s: unknown dtype %d", who, dtype);
    B200TRL_REQUIRE(n_rows >= 0 && vocab > 0 && row_stride >= vocab, B200TRL_E_INVALID,
                    "%s: bad shape rows=%lld vocab=%lld stride=%lld", who, (long long)n_rows, (long long)vocab,
                    (long long)row_stride);
    B200TRL_REQUIRE(inv_temperature > 0.f && std::isfinite(inv_temperature), B200TRL_E_INVALID,
                    "%s: inv_temperature must be positive and finite", who);
    a = K1Args{};
    a.logits = logits;
    a.n_rows = n_rows;
    a.vocab = vocab;
    a.row_stride = row_stride;
    B200TRL_REQUIRE(rows_per_batch >= 0 && (rows_per_batch == 0 || n_rows % rows_per_batch == 0), B200TRL_E_INVALID,
                    "%s: rows_per_batch %lld does not divide n_rows %lld", who, (long long)rows_per_batch,
                    (long long)n_rows);
    if (rows_per_batch > 0 && batch_stride == 0) batch_stride = rows_per_batch * row_stride;
    B200TRL_REQUIRE(rows_per_batch == 0 || batch_stride >= (rows_per_batch - 1) * row_stride + vocab, B200TRL_E_INVALID,
                    "%s: batch_stride %lld too small", who, (long long)batch_stride);
    if (rows_per_batch > 0 && batch_stride == rows_per_batch * row_stride) rows_per_batch = 0;  // flat after all
    a.rows_per_batch = rows_per_batch;
    a.batch_stride = rows_per_batch ? batch_stride : 0;
    a.ids = ids;
    a.inv_temp = inv_temperature;
    a.c = static_cast<float>(static_cast<double>(inv_temperature) * 1.4426950408889634);
    a.gmode = G_NONE;
    a.cfg.grad_scale = 1.f;
    a.grad_scale = 1.f;
    return B200TRL_OK;
}

// dlogits uses the same (batch, row) decomposition as the logits; a batched dlogits layout forces the logits to
// be addressed batched too (a flat logits tensor is a batched one with batch_stride = T * row_stride).
int set_dl_layout(K1Args& a, int64_t rows_per_batch, int64_t dl_batch_stride, const char* who) {
    if (rows_per_batch <= 0) {
        a.dl_batch_stride = 0;
        return 0;
    }
    if (dl_batch_stride == 0) dl_batch_stride = rows_per_batch * a.dl_row_stride;
    if (dl_batch_stride < (rows_per_batch - 1) * a.dl_row_stride + a.vocab) {
        set_error("%s: dl_batch_stride %lld too small", who, (long long)dl_batch_stride);
        return 1;
    }
    const bool dl_flat = (dl_batch_stride == rows_per_batch * a.dl_row_stride);
    if (a.rows_per_batch == 0 && dl_flat) {
        a.dl_batch_stride = 0;
        return 0;
    }
    if (a.rows_per_batch == 0) {  // logits flat, dlogits batched: address both batched
        a.rows_per_batch = rows_per_batch;
        a.batch_stride = rows_per_batch * a.row_stride;
    }
    a.dl_batch_stride = dl_batch_stride;
    return 0;
}

}  // namespace

extern "C" int b200trl_set_k1_path(int path) {
    if (path < B200TRL_K1_AUTO || path > B200TRL_K1_RESIDENT) return B200TRL_E_INVALID;
    return g_k1_path.exchange(path);
}

extern "C" int b200trl_set_skip_masked(int on) { return g_skip_masked.exchange(on ? 1 : 0); }

extern "C" int b200trl_k1_geometry(int64_t vocab, int mode, int32_t* out4) {
    B200TRL_REQUIRE(out4 != nullptr && mode >= 0 && mode <= 2, B200TRL_E_INVALID, "k1_geometry: bad arguments");
    k1_resident_geometry(vocab, mode, out4);
    return out4[0] ? B200TRL_OK : B200TRL_E_UNSUPPORTED;
}

extern "C" int b200trl_k1_set_trace(void* buffer, int64_t first_row) {
#ifdef B200TRL_K1_TRACE
    B200TRL_REQUIRE(first_row >= 0, B200TRL_E_INVALID, "k1_set_trace: bad arguments");
    g_trace_row0.store(static_cast<int>(first_row));
    g_trace.store(static_cast<unsigned long long*>(buffer));
    return B200TRL_OK;
#else
    (void)buffer;
    (void)first_row;
    set_error("k1_set_trace: this build has no trace hooks (make -C swh-trl_b200/csrc trace)");
    return B200TRL_E_UNSUPPORTED;
#endif
}

extern "C" int b200trl_logprob_entropy_fwd(const void* logits, int dtype, int64_t n_rows, int64_t vocab,
                                           int64_t row_stride, int64_t rows_per_batch, int64_t batch_stride,
                                           const int64_t* ids, float inv_temperature, float* logp, float* entropy,
                                           float* lse, b200trl_stream_t stream) {
    K1Args a;
    const int rc = fill_common(a, logits, dtype, n_rows, vocab, row_stride, rows_per_batch, batch_stride, ids,
                               inv_temperature, "logprob_entropy_fwd");
    if (rc) return rc;
    B200TRL_REQUIRE(logp, B200TRL_E_INVALID, "logprob_entropy_fwd: logp is null");
    a.logp = logp;
    a.entropy = entropy;
    a.lse = lse;
    return dispatch(a, dtype, as_stream(stream));
}

extern "C" int b200trl_masked_logprob_fwd(const void* logits, int dtype, int64_t n_rows, int64_t vocab,
                                          int64_t row_stride, int64_t rows_per_batch, int64_t batch_stride,
                                          const int64_t* ids, const uint8_t* row_mask, float inv_temperature,
                                          float* logp, float* entropy, float* lse, b200trl_stream_t stream) {
    K1Args a;
    const int rc = fill_common(a, logits, dtype, n_rows, vocab, row_stride, rows_per_batch, batch_stride, ids,
                               inv_temperature, "masked_logprob_fwd");
    if (rc) return rc;
    B200TRL_REQUIRE(logp && row_mask, B200TRL_E_INVALID, "masked_logprob_fwd: null pointer");
    a.row_mask = row_mask;
    a.logp = logp;
    a.entropy = entropy;
    a.lse = lse;
    return dispatch(a, dtype, as_stream(stream));
}

extern "C" int b200trl_logprob_bwd(const void* logits, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                                   int64_t rows_per_batch, int64_t batch_stride, const int64_t* ids,
                                   float inv_temperature, const float* lse, const float* g, void* dlogits,
                                   int64_t dl_row_stride, int64_t dl_batch_stride, b200trl_stream_t stream) {
    K1Args a;
    const int rc = fill_common(a, logits, dtype, n_rows, vocab, row_stride, rows_per_batch, batch_stride, ids,
                               inv_temperature, "logprob_bwd");
    if (rc) return rc;
    B200TRL_REQUIRE(lse && g && dlogits && dl_row_stride >= vocab, B200TRL_E_INVALID, "logprob_bwd: bad arguments");
    a.lse_in = lse;
    a.g = g;
    a.gmode = G_GIVEN;
    a.dlogits = dlogits;
    a.dl_row_stride = dl_row_stride;
    if (set_dl_layout(a, rows_per_batch, dl_batch_stride, "logprob_bwd")) return B200TRL_E_INVALID;
    return dispatch(a, dtype, as_stream(stream));
}

namespace {
constexpr int64_t kStepMaxClusters = 1024;  // >= 2 CTAs / SM x 148 SMs with room to spare
}

// workspace of b200trl_grpo_fused_step: [counter | cluster partials or K2's row partials] [total_count | row_count[B]]
static int64_t step_ws_main_bytes(int64_t B) {
    const int64_t in_kernel = 16 + kStepMaxClusters * 8 * static_cast<int64_t>(sizeof(float));
    return (std::max(in_kernel, b200trl_grpo_loss_workspace_bytes(B)) + 15) & ~int64_t(15);
}
extern "C" int64_t b200trl_grpo_fused_step_workspace_bytes(int64_t B) {
    return step_ws_main_bytes(B) + 16 + (B > 0 ? B : 0) * static_cast<int64_t>(sizeof(float));
}

static int grpo_fused_impl(const void* logits, int dtype, int64_t B, int64_t T, int64_t vocab, int64_t row_stride,
                           int64_t batch_stride, const int64_t* ids, const int32_t* mask, const float* advantages,
                           const float* old_logp, const float* ref_logp, const b200trl_grpo_cfg* cfg,
                           float inv_temperature, const float* row_count, const float* total_count, float* logp,
                           float* entropy, float* lse, void* dlogits, int64_t dl_row_stride, int64_t dl_batch_stride,
                           void* step_workspace, float* step_loss, float* step_metrics, b200trl_stream_t stream);

extern "C" int b200trl_grpo_fused_fwd_bwd(const void* logits, int dtype, int64_t B, int64_t T, int64_t vocab,
                                          int64_t row_stride, int64_t batch_stride, const int64_t* ids,
                                          const int32_t* mask, const float* advantages, const float* old_logp,
                                          const float* ref_logp, const b200trl_grpo_cfg* cfg, float inv_temperature,
                                          const float* row_count, const float* total_count, float* logp,
                                          float* entropy, float* lse, void* dlogits, int64_t dl_row_stride,
                                          int64_t dl_batch_stride, b200trl_stream_t stream) {
    return grpo_fused_impl(logits, dtype, B, T, vocab, row_stride, batch_stride, ids, mask, advantages, old_logp, ref_logp,
                           cfg, inv_temperature, row_count, total_count, logp, entropy, lse, dlogits, dl_row_stride,
                           dl_batch_stride, nullptr, nullptr, nullptr, stream);
}

extern "C" int b200trl_grpo_fused_step(const void* logits, int dtype, int64_t B, int64_t T, int64_t vocab,
                                       int64_t row_stride, int64_t batch_stride, const int64_t* ids, const int32_t* mask,
                                       const float* advantages, const float* old_logp, const float* ref_logp,
                                       const b200trl_grpo_cfg* cfg, float inv_temperature, const float* row_count,
                                       const float* total_count, float* logp, float* entropy, float* lse, void* dlogits,
                                       int64_t dl_row_stride, int64_t dl_batch_stride, void* workspace, float* loss,
                                       float* metrics, b200trl_stream_t stream) {
    B200TRL_REQUIRE(workspace && loss && metrics && entropy, B200TRL_E_INVALID, "grpo_fused_step: null pointer");
    return grpo_fused_impl(logits, dtype, B, T, vocab, row_stride, batch_stride, ids, mask, advantages, old_logp, ref_logp,
                           cfg, inv_temperature, row_count, total_count, logp, entropy, lse, dlogits, dl_row_stride,
                           dl_batch_stride, workspace, loss, metrics, stream);
}

static int grpo_fused_impl(const void* logits, int dtype, int64_t B, int64_t T, int64_t vocab, int64_t row_stride,
                           int64_t batch_stride, const int64_t* ids, const int32_t* mask, const float* advantages,
                           const float* old_logp, const float* ref_logp, const b200trl_grpo_cfg* cfg,
                           float inv_temperature, const float* row_count, const float* total_count, float* logp,
                           float* entropy, float* lse, void* dlogits, int64_t dl_row_stride, int64_t dl_batch_stride,
                           void* step_workspace, float* step_loss, float* step_metrics, b200trl_stream_t stream) {
    K1Args a;
    B200TRL_REQUIRE(B > 0 && T > 0, B200TRL_E_INVALID, "grpo_fused: bad shape");
    const int rc = fill_common(a, logits, dtype, B * T, vocab, row_stride, T, batch_stride, ids, inv_temperature,
                               "grpo_fused");
    if (rc) return rc;
    B200TRL_REQUIRE(mask && advantages && cfg && logp, B200TRL_E_INVALID, "grpo_fused: null pointer");
    B200TRL_REQUIRE((row_count && total_count) || (step_workspace && !row_count && !total_count), B200TRL_E_INVALID,
                    "grpo_fused: row_count / total_count must both be given (only b200trl_grpo_fused_step counts the mask "
                    "itself when both are null)");
    B200TRL_REQUIRE(cfg->loss_type >= 0 && cfg->loss_type <= 2, B200TRL_E_INVALID, "grpo_fused: unknown loss type %d",
                    cfg->loss_type);
    B200TRL_REQUIRE(cfg->is_level == B200TRL_IS_TOKEN || (cfg->is_level == B200TRL_IS_SEQUENCE && old_logp == nullptr),
                    B200TRL_E_UNSUPPORTED,
                    "grpo_fused: sequence-level importance sampling with old_logp needs the two-phase path");
    B200TRL_REQUIRE(cfg->beta == 0.f || ref_logp, B200TRL_E_INVALID, "grpo_fused: beta != 0 needs ref_logp");
    B200TRL_REQUIRE(!dlogits || dl_row_stride >= vocab, B200TRL_E_INVALID, "grpo_fused: bad dlogits stride");
    a.logp = logp;
    a.entropy = entropy;
    a.lse = lse;
    a.B = B;
    a.T = T;
    a.mask = mask;
    a.adv = advantages;
    a.old_lp = old_logp;
    a.ref_lp = (cfg->beta != 0.f) ? ref_logp : nullptr;
    a.row_count = row_count;
    a.total_count = total_count;
    a.cfg = *cfg;
    a.gmode = dlogits ? G_GRPO : G_NONE;
    a.skip_masked = (g_skip_masked.load() != 0 || cfg->skip_masked != 0) ? 1 : 0;
    a.dlogits = dlogits;
    a.dl_row_stride = dl_row_stride;
    if (dlogits && set_dl_layout(a, T, dl_batch_stride, "grpo_fused")) return B200TRL_E_INVALID;
    if (!step_workspace) return dispatch(a, dtype, as_stream(stream));
    // loss value + logged metrics in the same call: inside the resident kernel's pass when it takes the call with a
    // gradient, otherwise K2 right behind the pass (row kernel, forward-only evaluation)
    const bool in_kernel = dlogits != nullptr && takes_resident(a, dtype);
    if (!row_count) {
        // the mask statistics are part of the step: counted by the resident kernel's idle consumer warps while the
        // first chunks are in flight, or (row kernel, evaluation, very large batches) by the mask_stats kernel into
        // the tail of the workspace
        static const bool allow_count = []() {  // B200TRL_K1_COUNTMASK=0: always the mask_stats kernel (A/B runs)
            const char* v = getenv("B200TRL_K1_COUNTMASK");
            return !(v && atoi(v) == 0);
        }();
        if (allow_count && in_kernel && B <= 256 && B * T <= 131072) {
            a.count_mask = 1;
        } else {
            float* stats = reinterpret_cast<float*>(static_cast<char*>(step_workspace) + step_ws_main_bytes(B));
            const int rc0 = b200trl_mask_stats(mask, B, T, stats + 4, stats, stream);
            if (rc0) return rc0;
            row_count = stats + 4;
            total_count = stats;
            a.row_count = row_count;
            a.total_count = total_count;
        }
    }
    if (in_kernel) {
        a.step_ws = static_cast<float*>(step_workspace);
        a.step_loss = step_loss;
        a.step_metrics = step_metrics;
        return dispatch(a, dtype, as_stream(stream));
    }
    const int rc1 = dispatch(a, dtype, as_stream(stream));
    if (rc1) return rc1;
    b200trl_grpo_cfg c2 = *cfg;
    c2.grad_scale = 1.f;
    return b200trl_grpo_loss(logp, old_logp, a.ref_lp, advantages, mask, nullptr, entropy, B, T, &c2, row_count, total_count,
                             step_workspace, step_loss, step_metrics, nullptr, stream);
}

static int ppo_fused_impl(const void* logits, int dtype, int64_t mb, int64_t T, int64_t vocab, int64_t row_stride,
                          int64_t batch_stride, const int64_t* responses, const int64_t* sequence_lengths,
                          const float* old_logprobs, const float* advantages, float inv_temperature, float cliprange,
                          float grad_scale, float* new_logprobs, float* entropy, float* lse, void* dlogits,
                          int64_t dl_row_stride, int64_t dl_batch_stride, const float* returns, const float* values,
                          const float* vpred, float cliprange_value, float vf_coef, float* dvpred, void* workspace,
                          float* stats, b200trl_stream_t stream) {
    K1Args a;
    B200TRL_REQUIRE(mb > 0 && T > 0, B200TRL_E_INVALID, "ppo_fused: bad shape");
    const int rc = fill_common(a, logits, dtype, mb * T, vocab, row_stride, T, batch_stride, responses,
                               inv_temperature, "ppo_fused");
    if (rc) return rc;
    B200TRL_REQUIRE(sequence_lengths && old_logprobs && advantages && new_logprobs, B200TRL_E_INVALID,
                    "ppo_fused: null pointer");
    B200TRL_REQUIRE(!dlogits || dl_row_stride >= vocab, B200TRL_E_INVALID, "ppo_fused: bad dlogits stride");
    a.logp = new_logprobs;
    a.entropy = entropy;
    a.lse = lse;
    a.B = mb;
    a.T = T;
    a.seq_len = sequence_lengths;
    a.old_lp = old_logprobs;
    a.adv = advantages;
    a.clip_lo = static_cast<float>(1.0 - static_cast<double>(cliprange));
    a.clip_hi = static_cast<float>(1.0 + static_cast<double>(cliprange));
    a.grad_scale = grad_scale;
    a.gmode = G_PPO;  // also marks pad rows so that new_logprobs gets INVALID_LOGPROB there
    a.skip_masked = dlogits ? g_skip_masked.load() : 0;
    a.dlogits = dlogits;
    a.dl_row_stride = dl_row_stride;
    if (dlogits && set_dl_layout(a, T, dl_batch_stride, "ppo_fused")) return B200TRL_E_INVALID;
    if (!workspace) return dispatch(a, dtype, as_stream(stream));
    // the clipped losses, their statistics and d loss / d vpred in the same call: inside the resident kernel's pass
    // when it takes the call with a gradient, otherwise K2p right behind the pass
    if (dlogits != nullptr && takes_resident(a, dtype)) {
        a.step_ws = static_cast<float*>(workspace);
        a.step_metrics = stats;
        a.ppo_vpred = vpred;
        a.ppo_values = values;
        a.ppo_returns = returns;
        a.ppo_dvpred = dvpred;
        a.cliprange_value = cliprange_value;
        a.vf_coef = vf_coef;
        return dispatch(a, dtype, as_stream(stream));
    }
    const int rc1 = dispatch(a, dtype, as_stream(stream));
    if (rc1) return rc1;
    return b200trl_ppo_loss(new_logprobs, old_logprobs, advantages, returns, values, vpred, entropy, sequence_lengths, mb, T,
                            cliprange, cliprange_value, vf_coef, grad_scale, workspace, stats, dvpred, stream);
}

extern "C" int b200trl_ppo_fused_fwd_bwd(const void* logits, int dtype, int64_t mb, int64_t T, int64_t vocab,
                                         int64_t row_stride, int64_t batch_stride, const int64_t* responses,
                                         const int64_t* sequence_lengths, const float* old_logprobs,
                                         const float* advantages, float inv_temperature, float cliprange,
                                         float grad_scale, float* new_logprobs, float* entropy, float* lse,
                                         void* dlogits, int64_t dl_row_stride, int64_t dl_batch_stride,
                                         b200trl_stream_t stream) {
    return ppo_fused_impl(logits, dtype, mb, T, vocab, row_stride, batch_stride, responses, sequence_lengths, old_logprobs,
                          advantages, inv_temperature, cliprange, grad_scale, new_logprobs, entropy, lse, dlogits,
                          dl_row_stride, dl_batch_stride, nullptr, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, nullptr,
                          stream);
}

extern "C" int64_t b200trl_ppo_fused_step_workspace_bytes(int64_t mb) { return step_ws_main_bytes(mb); }

extern "C" int b200trl_ppo_fused_step(const void* logits, int dtype, int64_t mb, int64_t T, int64_t vocab,
                                      int64_t row_stride, int64_t batch_stride, const int64_t* responses,
                                      const int64_t* sequence_lengths, const float* old_logprobs, const float* advantages,
                                      const float* returns, const float* values, const float* vpred, float inv_temperature,
                                      float cliprange, float cliprange_value, float vf_coef, float grad_scale,
                                      float* new_logprobs, float* entropy, float* lse, void* dlogits,
                                      int64_t dl_row_stride, int64_t dl_batch_stride, float* dvpred, void* workspace,
                                      float* stats, b200trl_stream_t stream) {
    B200TRL_REQUIRE(returns && values && vpred && entropy && workspace && stats, B200TRL_E_INVALID,
                    "ppo_fused_step: null pointer");
    return ppo_fused_impl(logits, dtype, mb, T, vocab, row_stride, batch_stride, responses, sequence_lengths, old_logprobs,
                          advantages, inv_temperature, cliprange, grad_scale, new_logprobs, entropy, lse, dlogits,
                          dl_row_stride, dl_batch_stride, returns, values, vpred, cliprange_value, vf_coef, dvpred, workspace,
                          stats, stream);
}
