// Per-token loss math shared by the fused K1 kernels (inline gradient) and K2 (loss / metrics).
// GRPO: trl/trainer/grpo_trainer.py:2084-2137, 2158-2160.  PPO: trl/trainer/ppo_trainer.py:564-584.
#pragma once

#include "common.cuh"

namespace b200trl {

struct GrpoTok {
    float loss;   // per-token loss before masking / normalisation
    float dl;     // d(loss)/d(logp) through the surrogate only (importance-ratio path), KL excluded
    float dkl;    // beta * d(kl)/d(logp)
    float kl;     // k3 estimator (0 if beta == 0)
    float low;    // 1 if low-clipped
    float high;   // 1 if high-clipped
};

// Surrogate for an importance ratio r = exp(log_w) and its derivative wrt log_w.
// torch semantics kept: clamp passes gradient on [min, max] inclusive; min() splits ties 1/2 + 1/2.
__device__ __forceinline__ void grpo_surrogate(float log_w, float adv, const b200trl_grpo_cfg& cfg, float& loss,
                                               float& dloss_dlogw, float& low, float& high) {
    const float r = expf(log_w);                                    // coef_1 (:2113)
    const float c2 = fminf(fmaxf(r, cfg.clip_low), cfg.clip_high);  // coef_2 (:2114)
    float c1 = r;
    bool c1_pass = true;
    if (cfg.has_delta) {  // :2117-2118
        c1_pass = (r <= cfg.delta);
        c1 = fminf(r, cfg.delta);
    }
    const bool c2_pass = (r >= cfg.clip_low) && (r <= cfg.clip_high);
    const float l1 = c1 * adv, l2 = c2 * adv;  // :2120-2121
    const float d1 = c1_pass ? adv * r : 0.f;
    const float d2 = c2_pass ? adv * r : 0.f;
    const float dmin = (l1 < l2) ? d1 : ((l2 < l1) ? d2 : 0.5f * (d1 + d2));
    loss = -fminf(l1, l2);  // :2122
    dloss_dlogw = -dmin;
    low = (c1 < cfg.clip_low && adv < 0.f) ? 1.f : 0.f;    // :2158
    high = (c1 > cfg.clip_high && adv > 0.f) ? 1.f : 0.f;  // :2159
}

__device__ __forceinline__ void grpo_kl(float lp, float ref, float beta, float& kl, float& dkl) {
    const float d = ref - lp;  // :2087-2089
    const float e = expf(d);
    kl = e - d - 1.f;
    dkl = beta * (1.f - e);
}

// Token-level importance sampling (or old == lp.detach(), for which both levels coincide).
__device__ __forceinline__ GrpoTok grpo_token(float lp, bool has_old, float old, bool has_ref, float ref, float adv,
                                              float ent_keep, const b200trl_grpo_cfg& cfg) {
    GrpoTok t;
    float sl, sd;
    grpo_surrogate(has_old ? lp - old : 0.f, adv, cfg, sl, sd, t.low, t.high);
    t.loss = sl * ent_keep;  // :2123-2124
    t.dl = sd * ent_keep;
    t.kl = 0.f;
    t.dkl = 0.f;
    if (cfg.beta != 0.f && has_ref) {
        grpo_kl(lp, ref, cfg.beta, t.kl, t.dkl);
        t.loss += cfg.beta * t.kl;  // :2125-2126
    }
    return t;
}

// loss_type normalisation of a token in row b (:2130-2135), times the upstream grad scale
__device__ __forceinline__ float grpo_norm(const b200trl_grpo_cfg& cfg, float row_count, float total_count, float B) {
    switch (cfg.loss_type) {
        case B200TRL_LOSS_GRPO: return 1.f / (fmaxf(row_count, 1.f) * B);
        case B200TRL_LOSS_BNPO: return 1.f / fmaxf(total_count, 1.f);
        default: return 1.f / (B * cfg.max_completion_length);
    }
}

// PPO clipped policy term (ppo_trainer.py:578-583): returns max(pg1, pg2) and its derivative wrt new_logprob
__device__ __forceinline__ void ppo_policy(float new_lp, float old_lp, float adv, float lo, float hi, float& pg,
                                           float& dpg, float& clipped, float& ratio, float& diff) {
    diff = new_lp - old_lp;
    ratio = expf(diff);
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float pg1 = -adv * ratio, pg2 = -adv * rc;
    const float d1 = -adv * ratio;
    const float d2 = (ratio >= lo && ratio <= hi) ? -adv * ratio : 0.f;
    pg = fmaxf(pg1, pg2);
    dpg = (pg1 > pg2) ? d1 : ((pg2 > pg1) ? d2 : 0.5f * (d1 + d2));
    clipped = (pg2 > pg1) ? 1.f : 0.f;  // :589-591
}

}  // namespace b200trl
