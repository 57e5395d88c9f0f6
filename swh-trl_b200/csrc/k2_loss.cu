// K2: per-token loss bodies and their masked reductions, on [B,T] fp32 tensors (a few KB .. MB).
//   b200trl_mask_stats   grpo_trainer.py:2131,2133,2142 (mask row sums / total)
//   b200trl_grpo_loss    grpo_trainer.py:2084-2137 (loss) + :2139-2173 (local metric means)
//   b200trl_ppo_loss     ppo_trainer.py:564-605
// One CTA per sequence; row partials go to a workspace and the last CTA to finish folds them in row
// order (double accumulation), so results are run-to-run deterministic without a second launch.
#include "k1_args.cuh"

namespace b200trl {
namespace {

constexpr int kBlock = 256;
constexpr int kRowVals = 8;

struct Workspace {
    unsigned int counter;
    unsigned int pad[3];
    float rows[1];  // [B][kRowVals]
};

__global__ void __launch_bounds__(kBlock) mask_stats_kernel(const int32_t* __restrict__ mask, int64_t T,
                                                            float* __restrict__ row_count, float* total) {
    __shared__ float red[32];
    const int64_t b = blockIdx.x;
    float n[1] = {0.f};
    for (int64_t t = threadIdx.x; t < T; t += kBlock) n[0] += static_cast<float>(mask[b * T + t]);
    block_sum<1, kBlock>(n, red);
    if (threadIdx.x == 0) {
        row_count[b] = n[0];
        atomicAdd(total, n[0]);  // integer-valued addends: exact and order-independent below 2^24
    }
}

// true in exactly one CTA: the last one to get here; also resets the counter for the next launch
__device__ __forceinline__ bool last_block_done(unsigned int* counter) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counter, 1u);
        is_last = (prev == gridDim.x - 1);
        if (is_last) *counter = 0u;
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

struct GrpoLossArgs {
    const float* logp;
    const float* old_lp;
    const float* ref_lp;
    const float* adv;
    const int32_t* mask;
    const uint8_t* ent_mask;
    const float* entropy;
    int64_t B, T;
    b200trl_grpo_cfg cfg;
    const float* row_count;
    const float* total_count;
    Workspace* ws;
    float* loss;
    float* metrics;
    float* g;
};

__global__ void __launch_bounds__(kBlock) grpo_loss_kernel(const GrpoLossArgs a) {
    __shared__ float red[kRowVals * 32];
    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t base = b * a.T;
    const float len = a.row_count[b];
    const float total = a.total_count[0];
    const float Bf = static_cast<float>(a.B);
    const float norm = grpo_norm(a.cfg, len, total, Bf);
    const float adv = a.adv[b];
    const bool has_old = a.old_lp != nullptr, has_ref = a.ref_lp != nullptr;
    const bool seq = (a.cfg.is_level == B200TRL_IS_SEQUENCE);

    // sequence level: one importance weight per row from the masked mean log-ratio (:2102-2104)
    float seq_loss = 0.f, seq_d = 0.f, seq_low = 0.f, seq_high = 0.f;
    if (seq) {
        float s[1] = {0.f};
        if (has_old) {
            for (int64_t t = tid; t < a.T; t += kBlock)
                if (a.mask[base + t] != 0) s[0] += a.logp[base + t] - a.old_lp[base + t];
        }
        block_sum<1, kBlock>(s, red);
        grpo_surrogate(s[0] / fmaxf(len, 1.f), adv, a.cfg, seq_loss, seq_d, seq_low, seq_high);
    }

    // v: 0 loss, 1 kl, 2 entropy, 3 low, 4 high, 5 region, 6 sum(mask*ent_keep)
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t t = tid; t < a.T; t += kBlock) {
        const int64_t i = base + t;
        const float m = static_cast<float>(a.mask[i]);
        if (m == 0.f) {  // the loss ignores this token: never let its (possibly skipped / garbage) log-prob in
            if (!seq && a.g) a.g[i] = 0.f;
            continue;
        }
        const float keep = a.ent_mask ? static_cast<float>(a.ent_mask[i] != 0) : 1.f;
        const float lp = a.logp[i];
        float loss, kl = 0.f, dkl = 0.f, low, high, dl;
        if (seq) {
            loss = seq_loss * keep;
            dl = 0.f;
            low = seq_low;
            high = seq_high;
            if (a.cfg.beta != 0.f && has_ref) {
                grpo_kl(lp, a.ref_lp[i], a.cfg.beta, kl, dkl);
                loss += a.cfg.beta * kl;
            }
        } else {
            const GrpoTok tk = grpo_token(lp, has_old, has_old ? a.old_lp[i] : 0.f, has_ref, has_ref ? a.ref_lp[i] : 0.f,
                                          adv, keep, a.cfg);
            loss = tk.loss;
            kl = tk.kl;
            dkl = tk.dkl;
            dl = tk.dl;
            low = tk.low;
            high = tk.high;
            if (a.g) a.g[i] = m * norm * a.cfg.grad_scale * (dl + dkl);
        }
        v[0] += loss * m;
        v[1] += kl * m;
        if (a.entropy) v[2] += a.entropy[i] * m;
        v[3] += low * m;
        v[4] += high * m;
        v[5] += fmaxf(low, high) * m;
        v[6] += m * keep;
    }
    block_sum<7, kBlock>(v, red);

    if (seq && a.g) {
        // d loss / d logp_t' = norm * [ sum_t(mask*keep) * d(surrogate)/d(log_w) * mask_t' / max(len,1) + mask_t' * dkl_t' ]
        // (also when old == logp.detach(): the ratio is 1 but its gradient still flows through the row mean)
        const float coef = v[6] * seq_d / fmaxf(len, 1.f);
        for (int64_t t = tid; t < a.T; t += kBlock) {
            const int64_t i = base + t;
            const float m = static_cast<float>(a.mask[i]);
            if (m == 0.f) {
                a.g[i] = 0.f;
                continue;
            }
            float kl, dkl = 0.f;
            if (a.cfg.beta != 0.f && has_ref) grpo_kl(a.logp[i], a.ref_lp[i], a.cfg.beta, kl, dkl);
            a.g[i] = m * norm * a.cfg.grad_scale * (coef + dkl);
        }
    }

    if (tid == 0) {
        float* r = a.ws->rows + b * kRowVals;
        r[0] = v[0] * norm;
        r[1] = v[1];
        r[2] = v[2];
        // sequence level: clip indicators are [B,1] and averaged with a plain mean (:2144-2146)
        r[3] = seq ? seq_low : v[3];
        r[4] = seq ? seq_high : v[4];
        r[5] = seq ? fmaxf(seq_low, seq_high) : v[5];
        r[6] = 0.f;
        r[7] = 0.f;
    }
    if (!last_block_done(&a.ws->counter)) return;

    // ---- final fold over rows, in row order
    if (tid < 32) {
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int64_t r = tid; r < a.B; r += 32) {
            const volatile float* p = a.ws->rows + r * kRowVals;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[k] += static_cast<double>(p[k]);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
        if (tid == 0) {
            const double ntok = fmax(static_cast<double>(total), 1.0);
            const double clip_den = seq ? static_cast<double>(a.B) : ntok;
            a.loss[0] = static_cast<float>(acc[0]);
            a.metrics[B200TRL_M_LOSS] = static_cast<float>(acc[0]);
            a.metrics[B200TRL_M_KL] = static_cast<float>(acc[1] / ntok);
            a.metrics[B200TRL_M_ENTROPY] = static_cast<float>(acc[2] / ntok);
            a.metrics[B200TRL_M_CLIP_LOW] = static_cast<float>(acc[3] / clip_den);
            a.metrics[B200TRL_M_CLIP_HIGH] = static_cast<float>(acc[4] / clip_den);
            a.metrics[B200TRL_M_CLIP_REGION] = static_cast<float>(acc[5] / clip_den);
            a.metrics[B200TRL_M_TOKENS] = total;
            a.metrics[B200TRL_M_RESERVED] = 0.f;
        }
    }
}

struct PpoLossArgs {
    const float* new_lp;
    const float* old_lp;
    const float* adv;
    const float* returns;
    const float* values;
    const float* vpred;
    const float* entropy;
    const int64_t* seq_len;
    int64_t mb, T;
    float clip_lo, clip_hi, cliprange_value, vf_coef, grad_scale;
    Workspace* ws;
    float* stats;
    float* dvpred;
};

__global__ void __launch_bounds__(kBlock) ppo_loss_kernel(const PpoLossArgs a) {
    __shared__ float red[kRowVals * 32];
    __shared__ float s_cnt[2];
    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t base = b * a.T;
    if (tid < 32) {  // count(~padding_mask), count(~padding_mask_p1)  (ppo_trainer.py:501,505)
        float n0 = 0.f, n1 = 0.f;
        for (int64_t r = tid; r < a.mb; r += 32) {
            const int64_t len = a.seq_len[r];
            n0 += static_cast<float>(min(max(len + 1, (int64_t)0), a.T));
            n1 += static_cast<float>(min(max(len + 2, (int64_t)0), a.T));
        }
        n0 = warp_sum(n0);
        n1 = warp_sum(n1);
        if (tid == 0) {
            s_cnt[0] = n0;
            s_cnt[1] = n1;
        }
    }
    __syncthreads();
    const float n_pad = s_cnt[0], n_p1 = s_cnt[1];
    const int64_t len = a.seq_len[b];

    // v: 0 pg, 1 vf, 2 pg_clip, 3 vf_clip, 4 diff^2, 5 entropy, 6 ratio
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t t = tid; t < a.T; t += kBlock) {
        const int64_t i = base + t;
        const bool pad = t > len, pad1 = t > len + 1;
        const float dv = ppo_token_stats(pad ? 1.0f : a.new_lp[i], a.old_lp[i], a.adv[i], pad1 ? 0.f : a.vpred[i],
                                         a.values[i], a.returns[i], a.entropy ? a.entropy[i] : 0.f, pad, pad1, a.clip_lo,
                                         a.clip_hi, a.cliprange_value, v);
        if (a.dvpred) a.dvpred[i] = dv * (0.5f * a.vf_coef * a.grad_scale / n_p1);
    }
    block_sum<7, kBlock>(v, red);
    if (tid == 0) {
        float* r = a.ws->rows + b * kRowVals;
#pragma unroll
        for (int k = 0; k < 7; ++k) r[k] = v[k];
        r[7] = 0.f;
    }
    if (!last_block_done(&a.ws->counter)) return;
    if (tid < 32) {
        double acc[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int64_t r = tid; r < a.mb; r += 32) {
            const volatile float* p = a.ws->rows + r * kRowVals;
#pragma unroll
            for (int k = 0; k < 7; ++k) acc[k] += static_cast<double>(p[k]);
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) acc[k] = warp_sum(acc[k]);
        if (tid == 0) {
            const double n_all = static_cast<double>(a.mb) * static_cast<double>(a.T);
            const double pg_loss = acc[0] / n_pad;
            const double vf_loss = 0.5 * acc[1] / n_p1;
            a.stats[B200TRL_P_LOSS] = static_cast<float>(pg_loss + a.vf_coef * vf_loss);  // :584
            a.stats[B200TRL_P_PG_LOSS] = static_cast<float>(pg_loss);
            a.stats[B200TRL_P_VF_LOSS] = static_cast<float>(vf_loss);
            a.stats[B200TRL_P_PG_CLIPFRAC] = static_cast<float>(acc[2] / n_pad);
            a.stats[B200TRL_P_VF_CLIPFRAC] = static_cast<float>(acc[3] / n_p1);
            a.stats[B200TRL_P_APPROXKL] = static_cast<float>(0.5 * acc[4] / n_all);
            a.stats[B200TRL_P_ENTROPY] = static_cast<float>(acc[5] / n_all);
            a.stats[B200TRL_P_RATIO] = static_cast<float>(acc[6] / n_all);
        }
    }
}

}  // namespace
}  // namespace b200trl

using namespace b200trl;

extern "C" int b200trl_mask_stats(const int32_t* mask, int64_t B, int64_t T, float* row_count, float* total_count,
                                  b200trl_stream_t stream) {
    B200TRL_REQUIRE(mask && row_count && total_count, B200TRL_E_INVALID, "mask_stats: null pointer");
    B200TRL_REQUIRE(B > 0 && T > 0 && B <= 0x7fffffff, B200TRL_E_INVALID, "mask_stats: bad shape %lld x %lld",
                    (long long)B, (long long)T);
    cudaStream_t s = as_stream(stream);
    if (cudaMemsetAsync(total_count, 0, sizeof(float), s) != cudaSuccess) return check_launch("mask_stats memset");
    mask_stats_kernel<<<static_cast<unsigned>(B), kBlock, 0, s>>>(mask, T, row_count, total_count);
    return check_launch("mask_stats_kernel");
}

extern "C" int64_t b200trl_grpo_loss_workspace_bytes(int64_t B) {
    return static_cast<int64_t>(sizeof(Workspace)) + (B > 0 ? B : 0) * kRowVals * static_cast<int64_t>(sizeof(float));
}

extern "C" int b200trl_grpo_loss(const float* logp, const float* old_logp, const float* ref_logp,
                                 const float* advantages, const int32_t* mask, const uint8_t* ent_mask,
                                 const float* entropy, int64_t B, int64_t T, const b200trl_grpo_cfg* cfg,
                                 const float* row_count, const float* total_count, void* workspace, float* loss,
                                 float* metrics, float* g, b200trl_stream_t stream) {
    B200TRL_REQUIRE(logp && advantages && mask && cfg && row_count && total_count && workspace && loss && metrics,
                    B200TRL_E_INVALID, "grpo_loss: null pointer");
    B200TRL_REQUIRE(B > 0 && T > 0 && B <= 0x7fffffff, B200TRL_E_INVALID, "grpo_loss: bad shape");
    B200TRL_REQUIRE(cfg->loss_type >= 0 && cfg->loss_type <= 2, B200TRL_E_INVALID, "grpo_loss: unknown loss type %d",
                    cfg->loss_type);
    B200TRL_REQUIRE(cfg->is_level == 0 || cfg->is_level == 1, B200TRL_E_INVALID,
                    "grpo_loss: unknown importance sampling level %d", cfg->is_level);
    B200TRL_REQUIRE(cfg->beta == 0.f || ref_logp, B200TRL_E_INVALID, "grpo_loss: beta != 0 needs ref_logp");
    GrpoLossArgs a{logp, old_logp, ref_logp, advantages, mask, ent_mask, entropy, B, T, *cfg, row_count, total_count,
                   static_cast<Workspace*>(workspace), loss, metrics, g};
    grpo_loss_kernel<<<static_cast<unsigned>(B), kBlock, 0, as_stream(stream)>>>(a);
    return check_launch("grpo_loss_kernel");
}

extern "C" int b200trl_ppo_loss(const float* new_logprobs, const float* old_logprobs, const float* advantages,
                                const float* returns, const float* values, const float* vpred, const float* entropy,
                                const int64_t* sequence_lengths, int64_t mb, int64_t T, float cliprange,
                                float cliprange_value, float vf_coef, float grad_scale, void* workspace, float* stats,
                                float* dvpred, b200trl_stream_t stream) {
    B200TRL_REQUIRE(new_logprobs && old_logprobs && advantages && returns && values && vpred && sequence_lengths &&
                        workspace && stats,
                    B200TRL_E_INVALID, "ppo_loss: null pointer");
    B200TRL_REQUIRE(mb > 0 && T > 0 && mb <= 0x7fffffff, B200TRL_E_INVALID, "ppo_loss: bad shape");
    PpoLossArgs a{new_logprobs, old_logprobs, advantages, returns, values, vpred, entropy, sequence_lengths, mb, T,
                  static_cast<float>(1.0 - static_cast<double>(cliprange)),
                  static_cast<float>(1.0 + static_cast<double>(cliprange)), cliprange_value, vf_coef, grad_scale,
                  static_cast<Workspace*>(workspace), stats, dvpred};
    ppo_loss_kernel<<<static_cast<unsigned>(mb), kBlock, 0, as_stream(stream)>>>(a);
    return check_launch("ppo_loss_kernel");
}
