// Argument block shared by the two K1 implementations (row kernel, resident kernel) and the
// per-row epilogue that turns a finished log-prob into the scalar d(loss)/d(logp) of that token.
#pragma once

#include "token_math.cuh"

namespace b200trl {

enum GMode : int {
    G_NONE = 0,   // forward only (no dlogits)
    G_GIVEN = 1,  // g[row] supplied (two-phase path / plain autograd backward)
    G_GRPO = 2,   // GRPO loss gradient computed inline from per-token scalars
    G_PPO = 3     // PPO policy-loss gradient computed inline
};

struct K1Args {
    const void* logits;
    int64_t n_rows, vocab, row_stride;
    int64_t rows_per_batch;  // 0: flat (row r at r * row_stride); else row r = (b, t) at b * batch_stride + t * row_stride
    int64_t batch_stride;
    const int64_t* ids;
    const uint8_t* row_mask;  // forward-only, nullable: rows with row_mask[row] == 0 are not read (outputs 0)
    float c;         // inv_temperature * log2(e)
    float inv_temp;
    // forward outputs (nullable)
    float* logp;
    float* entropy;
    float* lse;
    // backward
    int gmode;
    void* dlogits;  // nullable
    int64_t dl_row_stride;
    int64_t dl_batch_stride;
    const float* lse_in;  // non-null => skip the forward pass (backward-only)
    const float* g;       // G_GIVEN
    // G_GRPO
    int64_t B, T;
    const int32_t* mask;
    const float* adv;
    const float* old_lp;
    const float* ref_lp;
    const float* row_count;
    const float* total_count;
    b200trl_grpo_cfg cfg;
    // G_PPO
    const int64_t* seq_len;
    float clip_lo, clip_hi, grad_scale;
    // fused modes only: rows whose loss weight is zero (completion_mask == 0 / PPO padding) are not read from HBM;
    // their dlogits are zeros and their logp / entropy / lse outputs are 0.  Opt-in (b200trl_set_skip_masked).
    int skip_masked;
    // G_GRPO, resident kernel only (b200trl_grpo_fused_step): the loss value and the logged metric sums are folded
    // into the same pass -- every cluster leader keeps running sums of its rows' terms, leaves them in step_ws and the
    // last one to finish adds the partials in cluster order (double) and writes loss[1] / metrics[8].  Null: off.
    int elem_f16;    // resident kernel: the 16-bit logits are fp16 (0: bf16); set by the dispatcher
    int count_mask;  // 1: row_count / total_count are null, the kernel counts `mask` itself (B <= 256 sequences)
    // G_PPO, resident kernel only (b200trl_ppo_fused_step): the clipped policy / value losses, their statistics and
    // d loss / d vpred (ppo_trainer.py:564-605) are produced by the same pass, like the GRPO sums below
    const float* ppo_vpred;
    const float* ppo_values;
    const float* ppo_returns;
    float* ppo_dvpred;  // nullable
    float cliprange_value, vf_coef;
    float* step_ws;  // [1 counter word + 3 pad][max clusters][8]
    float* step_loss;
    float* step_metrics;
    // diagnostics (b200trl_k1_set_trace, trace builds only): phase timestamps of the first 4 CTAs of the resident kernel
    unsigned long long* trace;
    int trace_row0;  // first traced row (per-CTA row counter)
};

// element offset of a row in the logits / dlogits tensors
__device__ __forceinline__ int64_t logits_offset(const K1Args& a, int64_t row) {
    if (a.rows_per_batch == 0) return row * a.row_stride;
    const int64_t b = row / a.rows_per_batch;
    return b * a.batch_stride + (row - b * a.rows_per_batch) * a.row_stride;
}
__device__ __forceinline__ int64_t dlogits_offset(const K1Args& a, int64_t row) {
    if (a.rows_per_batch == 0) return row * a.dl_row_stride;
    const int64_t b = row / a.rows_per_batch;
    return b * a.dl_batch_stride + (row - b * a.rows_per_batch) * a.dl_row_stride;
}

// true when the row's dlogits are known to be zero before reading it: the loss ignores it (fused GRPO / PPO
// modes) or its upstream gradient is exactly zero (backward of a masked token, e.g. DPO prompt positions)
__device__ __forceinline__ bool row_is_masked(const K1Args& a, int64_t row) {
    if (a.row_mask) return a.row_mask[row] == 0;
    if (a.gmode == G_GIVEN) return a.g[row] == 0.f;
    if (a.gmode == G_GRPO) return a.mask[row] == 0;
    if (a.gmode == G_PPO) return (row % a.T) > a.seq_len[row / a.T];
    return false;
}

// Number of non-pad positions sum_b min(len_b + 1 + shift, T) (ppo_trainer.py:501: pad = idx > len; :505: the mask
// shifted by one, shift = 1); warp-cooperative.
__device__ __forceinline__ float ppo_unpadded_count(const K1Args& a, int lane, int shift = 0) {
    float n = 0.f;
    for (int64_t b = lane; b < a.B; b += 32) {
        const int64_t len = a.seq_len[b];
        n += static_cast<float>(min(max(len + 1 + shift, (int64_t)0), a.T));
    }
    return warp_sum(n);
}

// One token's terms of the PPO micro-batch statistics (ppo_trainer.py:564-605), shared by K2p and the in-kernel sums of
// the resident K1 kernel.  v: 0 pg, 1 vf, 2 pg_clip, 3 vf_clip, 4 diff^2, 5 entropy, 6 ratio.  Returns d(loss)/d(vpred)
// without the 0.5 * vf_coef * grad_scale / n_p1 factor.
__device__ __forceinline__ float ppo_token_stats(float nlp, float old_lp, float adv, float vpred_raw, float val, float ret,
                                                 float entropy, bool pad, bool pad1, float clip_lo, float clip_hi,
                                                 float cliprange_value, float (&v)[7]) {
    float pg, dpg, clipped, ratio, diff;
    ppo_policy(pad ? 1.0f : nlp, old_lp, adv, clip_lo, clip_hi, pg, dpg, clipped, ratio, diff);  // :561-563 fill
    const float vp = pad1 ? 0.f : vpred_raw;  // :565
    const float lo = val - cliprange_value, hi = val + cliprange_value;
    const float vc = fminf(fmaxf(vp, lo), hi);  // :566-570
    const float e1 = vp - ret, e2 = vc - ret;
    const float vf1 = e1 * e1, vf2 = e2 * e2;  // :571-572
    const float keep = pad ? 0.f : 1.f, keep1 = pad1 ? 0.f : 1.f;
    v[0] += pg * keep;
    v[1] += fmaxf(vf1, vf2) * keep1;
    v[2] += clipped * keep;
    v[3] += (vf2 > vf1 ? 1.f : 0.f) * keep1;
    v[4] += diff * diff;  // unmasked mean (:594)
    v[5] += entropy;
    v[6] += ratio;
    const float d1 = 2.f * e1;
    const float d2 = (vp >= lo && vp <= hi) ? 2.f * e2 : 0.f;
    const float dmax = (vf1 > vf2) ? d1 : ((vf2 > vf1) ? d2 : 0.5f * (d1 + d2));
    return keep1 * dmax;
}

// Row scalars fetched early (latency hidden behind the streaming pass).
struct RowScalars {
    float x_sel;   // raw selected logit
    int64_t id;
    float aux0;    // G_GIVEN: g; G_GRPO: old; G_PPO: old
    float aux1;    // G_GRPO: ref
    float adv;
    float weight;  // mask * normalisation * grad_scale (0 => dlogits row is zero)
    float norm;    // G_GRPO: mask * normalisation (the token's weight in the loss VALUE)
    float pad;     // G_PPO: 1 if the position is padding
};

// Split in two so that a caller can software-pipeline the dependent pair (ids[row] -> logits[row, id]):
// everything addressed by the row alone ...  (`row_count` / `total_count`: where the mask statistics live -- the
// arguments' global arrays, or the resident kernel's shared-memory copy when it counted the mask itself)
__device__ __forceinline__ RowScalars load_row_scalars_direct(const K1Args& a, int64_t row, float ppo_count,
                                                              const float* row_count, const float* total_count) {
    RowScalars s;
    s.id = a.ids[row];
    s.x_sel = 0.f;
    s.aux0 = s.aux1 = s.adv = s.weight = s.pad = s.norm = 0.f;
    if (a.gmode == G_GIVEN) {
        s.aux0 = a.g[row];
        s.weight = 1.f;
    } else if (a.gmode == G_GRPO) {
        const int64_t b = row / a.T;
        const float m = static_cast<float>(a.mask[row]);
        s.adv = a.adv[b];
        if (a.old_lp) s.aux0 = a.old_lp[row];
        if (a.ref_lp) s.aux1 = a.ref_lp[row];
        s.norm = m * grpo_norm(a.cfg, row_count[b], total_count[0], static_cast<float>(a.B));
        s.weight = s.norm * a.cfg.grad_scale;
    } else if (a.gmode == G_PPO) {
        const int64_t b = row / a.T, t = row % a.T;
        s.pad = (t > a.seq_len[b]) ? 1.f : 0.f;
        s.adv = a.adv[row];
        s.aux0 = a.old_lp[row];
        s.weight = (1.f - s.pad) / ppo_count * a.grad_scale;
    }
    return s;
}
// ... and the selected logit, addressed through the id (NaN for an out-of-range id, as a gather would fault)
template <typename T>
__device__ __forceinline__ float load_selected_logit(const K1Args& a, int64_t row, int64_t id) {
    const T* base = reinterpret_cast<const T*>(a.logits) + logits_offset(a, row);
    return (id >= 0 && id < a.vocab) ? ElemTraits<T>::load(base + id) : __int_as_float(0x7fc00000);
}

__device__ __forceinline__ RowScalars load_row_scalars_direct(const K1Args& a, int64_t row, float ppo_count) {
    return load_row_scalars_direct(a, row, ppo_count, a.row_count, a.total_count);
}

template <typename T>
__device__ __forceinline__ RowScalars load_row_scalars(const K1Args& a, int64_t row, float ppo_count) {
    RowScalars s = load_row_scalars_direct(a, row, ppo_count);
    s.x_sel = load_selected_logit<T>(a, row, s.id);
    return s;
}

// d(loss)/d(logp) of this token, including mask, normalisation and the upstream scale
// `tok` (G_GRPO, optional): the token's loss terms for the in-kernel loss / metric sums; `tok->loss` etc. are left
// untouched for a token the loss ignores (s.norm == 0)
__device__ __forceinline__ float token_grad(const K1Args& a, const RowScalars& s, float logp, GrpoTok* tok = nullptr) {
    if (a.gmode == G_GIVEN) return s.aux0;
    if (a.gmode == G_GRPO) {
        if (s.norm == 0.f) return 0.f;
        const GrpoTok t = grpo_token(logp, a.old_lp != nullptr, s.aux0, a.ref_lp != nullptr, s.aux1, s.adv, 1.f, a.cfg);
        if (tok) *tok = t;
        return s.weight * (t.dl + t.dkl);
    }
    if (s.weight == 0.f) return 0.f;
    if (a.gmode == G_PPO) {
        float pg, dpg, clipped, ratio, diff;
        ppo_policy(logp, s.aux0, s.adv, a.clip_lo, a.clip_hi, pg, dpg, clipped, ratio, diff);
        return s.weight * dpg;
    }
    return 0.f;
}

int launch_k1_row(const K1Args& a, int dtype, cudaStream_t stream);
int launch_k1_resident(const K1Args& a, cudaStream_t stream);
bool k1_resident_supported(const K1Args& a, int dtype);
bool k1_resident_preferred(const K1Args& a, int dtype);
void k1_resident_geometry(int64_t vocab, int mode, int32_t out[4]);

}  // namespace b200trl
