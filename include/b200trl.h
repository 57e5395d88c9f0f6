/*
 * b200trl.h — C-ABI of libb200trl.so: the B200 (sm_100a) implementation of
 * TRL's per-token policy-loss hot path.
 *
 * The reference (shiwanghua/swh-trl, a fork of HF TRL 0.21.0.dev0) is pure
 * Python/torch and has no FFI of its own; each entry point below replaces the
 * torch-eager block cited next to it (paths relative to the reference root)
 * and is what a ctypes/cffi binding on the reference side would load
 * (INTEGRATION.md shows that stub).
 *
 * Conventions
 *   - plain device pointers and sizes only; no torch types;
 *   - every call is asynchronous on `stream`, allocates nothing, never
 *     synchronises and is CUDA-graph capturable; outputs and workspaces are
 *     caller-allocated;
 *   - return value: 0 = launched, <0 = B200TRL_E_* (nothing launched);
 *     b200trl_last_error() gives a thread-local message;
 *   - "rows" are logit-tokens: one (batch, position) pair = `vocab` logits.
 */
#ifndef B200TRL_H_
#define B200TRL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* b200trl_stream_t; /* == cudaStream_t */

enum b200trl_status {
    B200TRL_OK = 0,
    B200TRL_E_INVALID = -1,     /* bad argument (null pointer, size, enum)          */
    B200TRL_E_UNSUPPORTED = -2, /* dtype / alignment combination has no kernel      */
    B200TRL_E_LAUNCH = -3       /* cudaGetLastError() != cudaSuccess after launch   */
};

enum b200trl_dtype { B200TRL_BF16 = 0, B200TRL_F16 = 1, B200TRL_F32 = 2, B200TRL_F64 = 3 };

/* grpo_trainer.py:2130-2137 */
enum b200trl_loss_type { B200TRL_LOSS_GRPO = 0, B200TRL_LOSS_BNPO = 1, B200TRL_LOSS_DR_GRPO = 2 };
/* grpo_trainer.py:2100-2109 */
enum b200trl_is_level { B200TRL_IS_TOKEN = 0, B200TRL_IS_SEQUENCE = 1 };
/* ppo_trainer.py:511 */
enum b200trl_kl_estimator { B200TRL_KL_K1 = 0, B200TRL_KL_K3 = 1 };

/* which K1 implementation serves a call (b200trl_set_k1_path) */
enum b200trl_k1_path {
    B200TRL_K1_AUTO = 0,     /* smem-resident TMA kernel when bf16 + 16-byte aligned rows, else row kernel */
    B200TRL_K1_ROW = 1,      /* one CTA per row, plain vector loads, second pass served by L2            */
    B200TRL_K1_RESIDENT = 2  /* force the TMA-bulk, shared-memory-resident cluster kernel (error if n/a)  */
};

/* GRPO hyper-parameters as the loss body uses them (grpo_config.py:437-503). */
typedef struct b200trl_grpo_cfg {
    float beta;                  /* 0 disables the KL term (grpo_trainer.py:2085)                   */
    float clip_low;              /* (float)(1 - epsilon_low)  (grpo_trainer.py:2114)                */
    float clip_high;             /* (float)(1 + epsilon_high)                                       */
    float delta;                 /* upper clamp of coef_1, used iff has_delta (grpo_trainer.py:2117)*/
    int32_t has_delta;
    int32_t loss_type;           /* enum b200trl_loss_type                                          */
    int32_t is_level;            /* enum b200trl_is_level                                           */
    float max_completion_length; /* dr_grpo denominator (grpo_trainer.py:2135)                      */
    float grad_scale;            /* upstream d(loss) known a priori (1/grad_accum, AMP scale)       */
    int32_t skip_masked;         /* fused passes: per-call form of b200trl_set_skip_masked -- rows with
                                    completion_mask == 0 are not read (their dlogits are zeros, their
                                    per-token outputs 0); loss, metrics and gradients are unchanged    */
} b200trl_grpo_cfg;

/* metrics[] layout written by b200trl_grpo_loss (grpo_trainer.py:2150-2172, local means) */
enum b200trl_grpo_metric {
    B200TRL_M_LOSS = 0,
    B200TRL_M_KL = 1,          /* masked batch mean of the k3 KL (0 if beta == 0)        */
    B200TRL_M_ENTROPY = 2,     /* masked batch mean entropy (0 if no entropy given)      */
    B200TRL_M_CLIP_LOW = 3,
    B200TRL_M_CLIP_HIGH = 4,
    B200TRL_M_CLIP_REGION = 5,
    B200TRL_M_TOKENS = 6,      /* completion_mask.sum() (before clamp)                   */
    B200TRL_M_RESERVED = 7,
    B200TRL_GRPO_NUM_METRICS = 8
};

/* stats[] layout written by b200trl_ppo_loss (ppo_trainer.py:573-605) */
enum b200trl_ppo_stat {
    B200TRL_P_LOSS = 0,
    B200TRL_P_PG_LOSS = 1,
    B200TRL_P_VF_LOSS = 2,
    B200TRL_P_PG_CLIPFRAC = 3,
    B200TRL_P_VF_CLIPFRAC = 4,
    B200TRL_P_APPROXKL = 5,
    B200TRL_P_ENTROPY = 6,
    B200TRL_P_RATIO = 7,
    B200TRL_PPO_NUM_STATS = 8
};

/* stats[] layout written by b200trl_rloo_loss (rloo_trainer.py:486-507) */
enum b200trl_rloo_stat {
    B200TRL_R_LOSS = 0,        /* pg_loss = mean_b max(-A r, -A clamp(r)), r = exp(sum_t new - sum_t old) */
    B200TRL_R_PG_CLIPFRAC = 1,
    B200TRL_R_APPROXKL = 2,
    B200TRL_R_ENTROPY = 3,
    B200TRL_R_RATIO = 4,       /* mean token-level ratio (new_ratio, :476) */
    B200TRL_RLOO_NUM_STATS = 8
};

int b200trl_version(void);
const char* b200trl_last_error(void);
int b200trl_set_k1_path(int path); /* returns the previous setting */
/* Opt-in: in the fused GRPO / PPO passes, rows the loss ignores (completion_mask == 0, PPO padding) are not read
 * from HBM; their dlogits are zeros and their logp / entropy / lse outputs are 0 (PPO: 1.0).  Loss, metrics and
 * gradients are unchanged; only the per-token outputs at masked positions differ from the reference (which
 * computes and then discards them).  Returns the previous setting. */
int b200trl_set_skip_masked(int on);
/* Which shape the resident K1 kernel takes for a bf16 vocabulary (host-only, no GPU needed): out4 = {consumer threads
 * per CTA, cluster size, ring slots, chunk bytes}; mode 0 forward-only, 1 backward-only, 2 fused.  Returns
 * B200TRL_E_UNSUPPORTED (out4 zeroed) when the call would go to the row kernel.  See DESIGN.md §3. */
int b200trl_k1_geometry(int64_t vocab, int mode, int32_t* out4);
/* Diagnostics, trace builds only (`make -C swh-trl_b200/csrc trace` -> lib/libb200trl_trace.so; the production
 * library returns B200TRL_E_UNSUPPORTED): while `buffer` (device memory, 4 * 3 * 169 uint64, zeroed by the caller) is
 * set, the resident K1 kernel logs phase timestamps (tag << 56 | row << 40 | chunk << 32 | clock32) of three roles
 * (consumer warp 0, reducer, DMA) of its first four CTAs for 8 rows per CTA starting at its `first_row`-th row; each
 * role's block starts with its event count.  NULL switches tracing off.  See tools/k1_trace.py. */
int b200trl_k1_set_trace(void* buffer, int64_t first_row);

/* ---- K1: selective_log_softmax + entropy_from_logits, one pass ------------------------------
 * Replaces trl/trainer/utils.py:1430-1462 and :1465-1490 (and the division by the temperature,
 * grpo_trainer.py:1258 / ppo_trainer.py:450,559, folded in as `inv_temperature`).
 * logits: [n_rows, vocab] with `row_stride` elements between rows.  With rows_per_batch = T > 0 the
 * rows are addressed as (b, t) = (r / T, r % T) at b * batch_stride + t * row_stride, i.e. the
 * `logits[:, :-1][:, -T:]` view of the model output is read in place and the slice copy at
 * grpo_trainer.py:1252-1254 is not needed (rows_per_batch = 0: flat; batch_stride = 0: dense).
 * logp/entropy/lse: fp32 [n_rows]; entropy and lse may be NULL. */
int b200trl_logprob_entropy_fwd(const void* logits, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                                int64_t rows_per_batch, int64_t batch_stride, const int64_t* ids,
                                float inv_temperature, float* logp, float* entropy, float* lse,
                                b200trl_stream_t stream);

/* The same forward with a row mask (uint8 [n_rows], a torch.bool tensor's storage): rows with row_mask == 0 are NOT
 * read and get logp = entropy = lse = 0.  This is what the DPO-family consumers compute around the call
 * (`labels[~loss_mask] = 0; per_token_logps = selective_log_softmax(logits, labels); per_token_logps[~loss_mask] = 0`,
 * trl/trainer/dpo_trainer.py:1557-1559; likewise kto / bco / cpo / orpo), minus the reads of the prompt rows. */
int b200trl_masked_logprob_fwd(const void* logits, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                               int64_t rows_per_batch, int64_t batch_stride, const int64_t* ids,
                               const uint8_t* row_mask, float inv_temperature, float* logp, float* entropy, float* lse,
                               b200trl_stream_t stream);

/* Backward of the gather-log-softmax: dlogits[r,v] = g[r]*inv_T*(1[v==ids[r]] - exp(x[r,v]*inv_T - lse[r])),
 * written in the logits dtype.  What autograd produces for utils.py:1449-1461 + grpo_trainer.py:1258. */
int b200trl_logprob_bwd(const void* logits, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                        int64_t rows_per_batch, int64_t batch_stride, const int64_t* ids, float inv_temperature,
                        const float* lse, const float* g, void* dlogits, int64_t dl_row_stride,
                        int64_t dl_batch_stride, b200trl_stream_t stream);

/* completion_mask row sums and total (grpo_trainer.py:2131,2133,2142); row_count fp32 [B], total fp32 [1]. */
int b200trl_mask_stats(const int32_t* mask, int64_t B, int64_t T, float* row_count, float* total_count,
                       b200trl_stream_t stream);

/* ---- K1+K2 fused: log-prob, entropy AND dlogits of the GRPO loss in one pass over the logits ---
 * Replaces grpo_trainer.py:2067 (+ :1252-1268) and the autograd backward of :2084-2137 down to the
 * logits.  Valid when d(loss)/d(logp) of a token is known from that token alone: token-level
 * importance sampling, or old_logp == NULL (ratio == 1 for either level), and no entropy mask.
 * old_logp / ref_logp may be NULL (ref must be given iff cfg->beta != 0).  row_count/total_count
 * come from b200trl_mask_stats.  dlogits == NULL gives the forward only.  batch_stride /
 * dl_batch_stride (0 = dense) let logits and dlogits live inside the model's [B, L, V] tensors. */
int b200trl_grpo_fused_fwd_bwd(const void* logits, int dtype, int64_t B, int64_t T, int64_t vocab, int64_t row_stride,
                               int64_t batch_stride, const int64_t* ids, const int32_t* mask,
                               const float* advantages, const float* old_logp, const float* ref_logp,
                               const b200trl_grpo_cfg* cfg, float inv_temperature, const float* row_count,
                               const float* total_count, float* logp, float* entropy, float* lse, void* dlogits,
                               int64_t dl_row_stride, int64_t dl_batch_stride, b200trl_stream_t stream);

/* The same pass PLUS the loss value and the logged metric means of grpo_trainer.py:2130-2135, 2139-2173 in the same
 * call (what b200trl_grpo_loss would compute from the log-probs with ent_mask == NULL, g == NULL): on the resident
 * kernel with dlogits != NULL every cluster keeps running sums of its rows' loss terms and the last cluster to finish
 * folds them in cluster order (double accumulation, deterministic for a given GPU) -- no second launch; otherwise
 * (row kernel, forward-only evaluation) K2 is launched right behind the pass.  `workspace`:
 * b200trl_grpo_fused_step_workspace_bytes(B) bytes, zeroed once by the caller (the counter resets itself); it is
 * in use until the call's work on `stream` has finished.  loss fp32 [1], metrics fp32 [B200TRL_NUM_GRPO_METRICS];
 * entropy must be given.  row_count and total_count may BOTH be NULL: the call then counts the completion mask itself
 * (grpo_trainer.py:2131,2133,2142) -- inside the resident kernel, by its consumer warps while the first chunks are in
 * flight (B <= 256 sequences, B*T <= 131072 tokens), else with the mask_stats kernel into the workspace's tail. */
int64_t b200trl_grpo_fused_step_workspace_bytes(int64_t B);
int b200trl_grpo_fused_step(const void* logits, int dtype, int64_t B, int64_t T, int64_t vocab, int64_t row_stride,
                            int64_t batch_stride, const int64_t* ids, const int32_t* mask, const float* advantages,
                            const float* old_logp, const float* ref_logp, const b200trl_grpo_cfg* cfg,
                            float inv_temperature, const float* row_count, const float* total_count, float* logp,
                            float* entropy, float* lse, void* dlogits, int64_t dl_row_stride, int64_t dl_batch_stride,
                            void* workspace, float* loss, float* metrics, b200trl_stream_t stream);

/* ---- K2: GRPO loss body + metrics (+ per-token d(loss)/d(logp)) --------------------------------
 * Replaces grpo_trainer.py:2084-2137 (loss) and :2139-2173 (local metric means).  All [B,T]
 * inputs fp32 row-major; ent_mask (uint8, from b200trl_entropy_quantile_mask) and entropy may be
 * NULL.  workspace: >= b200trl_grpo_loss_workspace_bytes(B) bytes, zero-initialised once (the
 * kernel leaves it zeroed).  loss: fp32 [1]; metrics: fp32 [B200TRL_GRPO_NUM_METRICS];
 * g: fp32 [B,T] or NULL. */
int64_t b200trl_grpo_loss_workspace_bytes(int64_t B);
int b200trl_grpo_loss(const float* logp, const float* old_logp, const float* ref_logp, const float* advantages,
                      const int32_t* mask, const uint8_t* ent_mask, const float* entropy, int64_t B, int64_t T,
                      const b200trl_grpo_cfg* cfg, const float* row_count, const float* total_count, void* workspace,
                      float* loss, float* metrics, float* g, b200trl_stream_t stream);

/* ---- a-4: get_high_entropy_mask (grpo_trainer.py:341-364) ------------------------------------
 * out_mask[i] = mask[i] && entropies[i]*mask[i] >= quantile(entropies[mask], threshold) with torch's
 * linear interpolation; all zero if the mask is empty.  workspace >= ..._workspace_bytes(n), any
 * contents.  out_threshold (fp32 [1]) may be NULL.  One cooperative launch (multi-CTA radix select). */
int64_t b200trl_entropy_quantile_workspace_bytes(int64_t n);
int b200trl_entropy_quantile_mask(const float* entropies, const int32_t* mask, int64_t n, float threshold,
                                  void* workspace, uint8_t* out_mask, float* out_threshold, b200trl_stream_t stream);

/* ---- K3: group-relative advantages (grpo_trainer.py:1917-1938) -------------------------------
 * rewards_per_func: gathered, rank-major fp32 [B_global, n_funcs] (NaN = not applicable).
 * Outputs: rewards [B_global], adv_all [B_global], adv_local [local_count] =
 * adv_all[local_offset : local_offset+local_count], mean/std fp32 [B_global/G], is_std_zero uint8. */
int b200trl_group_advantages(const float* rewards_per_func, const float* weights, int64_t B_global, int64_t n_funcs,
                             int64_t G, int scale_rewards, int64_t local_offset, int64_t local_count, float* rewards,
                             float* adv_all, float* adv_local, float* mean, float* std, uint8_t* is_std_zero,
                             b200trl_stream_t stream);

/* ---- f-4: the logging block of GRPOTrainer._generate_and_score_completions (grpo_trainer.py:1942-1972) -----------
 * Three gathers + ~13 .item() host syncs (+ 2 per reward function) become ONE gather of a packed int64 vector, this ONE
 * launch and ONE device->host read.  packed: the all-gathered vector, rank-major [world][1 + 2*b_local] =
 * {sum(attention_mask), completion_lengths[b_local], terminated_with_eos[b_local]}; rewards_per_func: gathered fp32
 * [world*b_local, n_funcs] (NaN = not applicable); mean / std / is_std_zero: per group, from b200trl_group_advantages.
 * out: fp64 [12 + 2*n_funcs] = {num_input_tokens, completions/mean_length, min_length, max_length, clipped_ratio,
 * mean_terminated_length, min_terminated_length, max_terminated_length (0 when nothing terminated, :1956-1957),
 * reward, reward_std, frac_reward_zero_std, 0, then (nanmean, nanstd) per reward function (:196-211, :1961-1965)}. */
int b200trl_generation_stats(const int64_t* packed, int64_t world, int64_t b_local, const float* rewards_per_func,
                             int64_t n_funcs, const float* mean_grouped, const float* std_grouped,
                             const uint8_t* is_std_zero, int64_t n_groups, double* out, b200trl_stream_t stream);

/* ---- K4: PPO KL reward + score + whitening + reverse GAE (ppo_trainer.py:500-535) ------------
 * logprobs/ref_logprobs/values: raw fp32 [B,T] (pads filled here as :500-506 does); scores [B];
 * sequence_lengths int64 [B].  Outputs fp32 [B,T]: rewards, advantages (whitened, pads 0), returns;
 * optional filled copies logprobs_f/ref_logprobs_f/values_f (NULL to skip).
 * workspace >= b200trl_ppo_gae_workspace_bytes(B,T), any contents.  One cooperative launch. */
int64_t b200trl_ppo_gae_workspace_bytes(int64_t B, int64_t T);
int b200trl_ppo_rewards_gae(const float* logprobs, const float* ref_logprobs, const float* values, const float* scores,
                            const int64_t* sequence_lengths, int64_t B, int64_t T, float kl_coef, int kl_estimator,
                            float gamma, float lam, int whiten_rewards, void* workspace, float* rewards,
                            float* advantages, float* returns, float* logprobs_f, float* ref_logprobs_f,
                            float* values_f, b200trl_stream_t stream);

/* ---- PPO policy/value loss (ppo_trainer.py:557-605) --------------------------------------------
 * fused pass: new_logprobs (pads = 1.0), entropy and dlogits of pg_loss for a micro-batch [mb,T,V];
 * the token weight is (1 - pad)/count(~pad) * grad_scale.  dlogits may be NULL. */
int b200trl_ppo_fused_fwd_bwd(const void* logits, int dtype, int64_t mb, int64_t T, int64_t vocab, int64_t row_stride,
                              int64_t batch_stride, const int64_t* responses, const int64_t* sequence_lengths,
                              const float* old_logprobs, const float* advantages, float inv_temperature,
                              float cliprange, float grad_scale, float* new_logprobs, float* entropy, float* lse,
                              void* dlogits, int64_t dl_row_stride, int64_t dl_batch_stride,
                              b200trl_stream_t stream);
/* loss + stats from new_logprobs (as written above) and the value head; dvpred [mb,T] may be NULL.
 * workspace as for b200trl_grpo_loss with B = mb. */
int b200trl_ppo_loss(const float* new_logprobs, const float* old_logprobs, const float* advantages,
                     const float* returns, const float* values, const float* vpred, const float* entropy,
                     const int64_t* sequence_lengths, int64_t mb, int64_t T, float cliprange, float cliprange_value,
                     float vf_coef, float grad_scale, void* workspace, float* stats, float* dvpred,
                     b200trl_stream_t stream);

/* The PPO micro-batch in ONE call: b200trl_ppo_fused_fwd_bwd plus everything b200trl_ppo_loss computes from its
 * outputs (the clipped policy / value losses, the seven statistics and d loss / d vpred, ppo_trainer.py:564-605).  On
 * the resident kernel with dlogits != NULL the per-token terms are summed inside the pass (cluster partials, folded by
 * the last cluster in cluster order, double) -- no second launch; otherwise (row kernel, evaluation) K2p is launched
 * right behind the pass.  workspace: b200trl_ppo_fused_step_workspace_bytes(mb) bytes, zeroed once (the counter resets
 * itself).  dvpred (fp32 [mb,T]) may be NULL; entropy must be given. */
int64_t b200trl_ppo_fused_step_workspace_bytes(int64_t mb);
int b200trl_ppo_fused_step(const void* logits, int dtype, int64_t mb, int64_t T, int64_t vocab, int64_t row_stride,
                           int64_t batch_stride, const int64_t* responses, const int64_t* sequence_lengths,
                           const float* old_logprobs, const float* advantages, const float* returns, const float* values,
                           const float* vpred, float inv_temperature, float cliprange, float cliprange_value,
                           float vf_coef, float grad_scale, float* new_logprobs, float* entropy, float* lse,
                           void* dlogits, int64_t dl_row_stride, int64_t dl_batch_stride, float* dvpred, void* workspace,
                           float* stats, b200trl_stream_t stream);

/* ---- RLOO ---------------------------------------------------------------------
 * rewards + leave-one-out advantages, trl/trainer/rloo_trainer.py:397-441.  logprobs/ref_logprobs raw fp32 [B,T]
 * (pads filled here), scores [B], sequence_lengths int64 [B]; sample i of prompt p is row i * (B / rloo_k) + p.
 * Outputs fp32 [B]: advantages, rlhf_reward, non_score_reward; optional filled log-probs. */
int b200trl_rloo_rewards_advantages(const float* logprobs, const float* ref_logprobs, const float* scores,
                                    const int64_t* sequence_lengths, int64_t B, int64_t T, float kl_coef,
                                    int64_t rloo_k, int normalize_reward, float reward_clip_range,
                                    int normalize_advantage, int token_level_kl, float* advantages,
                                    float* rlhf_reward, float* non_score_reward, float* logprobs_f,
                                    float* ref_logprobs_f, b200trl_stream_t stream);
/* sequence-ratio clipped loss + stats (rloo_trainer.py:476-507) from new log-probs; g fp32 [mb,T] (nullable) is
 * d(loss)/d(new_logprob) for b200trl_logprob_bwd.  workspace as for b200trl_grpo_loss with B = mb. */
int b200trl_rloo_loss(const float* new_logprobs, const float* old_logprobs, const float* advantages,
                      const float* entropy, const int64_t* sequence_lengths, int64_t mb, int64_t T, float cliprange,
                      float grad_scale, void* workspace, float* stats, float* g, b200trl_stream_t stream);

/* ---- K5 (forward): lm_head GEMM fused with the log-softmax statistics on tcgen05 / TMEM ----------------------
 * logp / entropy / lse of (hidden @ weight^T) * inv_temperature without ever materialising the [n_rows, vocab]
 * logits (what `_get_per_token_logps_and_entropies` needs for the no-grad old / ref passes when the
 * Liger path is on, grpo_trainer.py:1163-1203 + :1261-1267).  hidden: bf16 [n_rows, hidden_size]; weight: bf16
 * [vocab, hidden_size] (lm_head.weight); strides in elements, multiples of 8.  workspace >=
 * b200trl_fused_linear_workspace_bytes(n_rows, vocab), any contents.  Two launches (GEMM+stats, merge). */
int64_t b200trl_fused_linear_workspace_bytes(int64_t n_rows, int64_t vocab);
int b200trl_fused_linear_logprob_fwd(const void* hidden, int64_t hidden_row_stride, const void* weight,
                                     int64_t weight_row_stride, int64_t n_rows, int64_t hidden_size, int64_t vocab,
                                     const int64_t* ids, float inv_temperature, void* workspace, float* logp,
                                     float* entropy, float* lse, b200trl_stream_t stream);

/* ---- K7: the three contractions of the Liger seam on tcgen05 (CTA-pair MMA, TMA, TMEM accumulators) -----------
 * D[M,N] = A[M,K] * B[N,K]^T, bf16 operands, fp32 accumulation: what `LigerFusedLinearGRPOLoss` contracts per chunk
 * (grpo_trainer.py:2005-2045): logits_c = hidden_c W^T, dH_c = dlogits_c W, dW += dlogits_c^T hidden_c.
 * a_layout / b_layout: 0 = the operand is stored [rows, K] (K contiguous), 1 = stored [K, rows] (rows contiguous);
 * lda / ldb / ldd: elements between stored rows (multiples of 8).  out_kind B200TRL_TC_OUT_BF16: out = bf16 [M, ldd],
 * D (+ bias[N], bf16, nullable) (+ addend fp32 [M, ld_addend], nullable) rounded once; B200TRL_TC_OUT_F32_ACC: out =
 * fp32 [M, ldd], out += D; B200TRL_TC_OUT_F32: out = D.
 * Supported (a_layout, b_layout, out_kind): (0,0,BF16) (0,1,BF16) (1,1,BF16) (1,1,F32_ACC) (1,1,F32).  m_fastest: work order, 1 = clusters
 * running at the same time share the B tile, 0 = the A tile (pick the operand that does not fit L2).
 * workspace (nullable, b200trl_tc_gemm_workspace_bytes, any contents): fp32 scratch for the bf16 output.  With it, a
 * contraction that has too few output tiles to fill the 74 CTA pairs (dH: 16 x 14 tiles, K = 152064) is split along
 * K; the slices' partial products land in planes of the scratch and are added in slice order (deterministic). */
enum { B200TRL_TC_OUT_BF16 = 1, B200TRL_TC_OUT_F32_ACC = 2, B200TRL_TC_OUT_F32 = 3 };
int64_t b200trl_tc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int out_kind);
int b200trl_tc_gemm(const void* A, int a_layout, int64_t lda, const void* B, int b_layout, int64_t ldb, int64_t M,
                    int64_t N, int64_t K, int out_kind, void* out, int64_t ldd, const void* bias, const float* addend,
                    int64_t ld_addend, int m_fastest, void* workspace, int64_t workspace_bytes,
                    b200trl_stream_t stream);
/* Which implementation the seam's three GEMMs use: bit 0 = logits, bit 1 = dH, bit 2 = dW; a set bit = the tcgen05
 * kernel above (default: all three), a clear bit = cuBLASLt (A/B baseline).  Returns the previous mask; mask < 0
 * only queries. */
int b200trl_set_seam_gemm_mask(int mask);

/* ---- a-13: the reference's operator seam as ONE call (grpo_trainer.py:870-886 ctor, :2005-2045 call) ----------
 * What `self.liger_grpo_loss(_input, lin_weight, selected_token_ids, attention_mask, advantages, bias,
 * old_per_token_logps, ref_per_token_logps)` computes, forward AND backward, without ever holding the [B,T,V]
 * logits: per chunk of `chunk_seqs` whole sequences  logits_c = hidden_c W^T (+ bias)  (this library's CTA-pair
 * tcgen05 GEMM, bf16 in, fp32 accumulate in TMEM)  ->  K1 resident kernel in place (logits_c becomes dlogits_c;
 * log-probs, entropies)  ->  dH_c = dlogits_c W  and  dW += dlogits_c^T hidden_c  (same kernel family; dW is
 * accumulated in fp32 by TMA reduce-add); then K2 gives the loss and metrics.  cuBLASLt can be selected per GEMM
 * as an A/B baseline (b200trl_set_seam_gemm_mask / B200TRL_SEAM_GEMM, default 7 = all ours); it is bound with
 * dlopen only if selected.
 * hidden bf16 [B,T,H]; weight bf16 [V,H]; bias bf16 [V] or NULL; H, V multiples of 8.
 * Outputs: logp, entropy fp32 [B,T]; loss fp32 [1]; metrics fp32 [B200TRL_GRPO_NUM_METRICS] (metrics[B200TRL_M_KL]
 * and [B200TRL_M_CLIP_REGION] are what grpo_trainer.py:2038-2039 logs); dhidden bf16 [B,T,H], dweight fp32 [V,H]
 * (the accumulator over the chunks: written by the first chunk, no pre-zeroing), dweight_bf16 bf16 [V,H] (the final
 * dW rounded once to the weight's dtype — with the tcgen05 dW GEMM the last chunk's epilogue emits it as
 * bf16(dweight + D) and dweight itself is then left one chunk short; dweight may be NULL when B <= chunk_seqs),
 * dbias fp32 [V] — each nullable; all NULL = forward only.  Gradients are of loss * cfg->grad_scale.
 * workspace >= b200trl_fused_linear_grpo_workspace_bytes(B, T, H, V, chunk_seqs) bytes, any contents. */
int64_t b200trl_fused_linear_grpo_workspace_bytes(int64_t B, int64_t T, int64_t H, int64_t V, int64_t chunk_seqs);
int b200trl_fused_linear_grpo(const void* hidden, const void* weight, const void* bias, int64_t B, int64_t T,
                              int64_t H, int64_t V, const int64_t* ids, const int32_t* mask, const float* advantages,
                              const float* old_logp, const float* ref_logp, const b200trl_grpo_cfg* cfg,
                              float inv_temperature, int64_t chunk_seqs, void* workspace, float* logp, float* entropy,
                              float* loss, float* metrics, void* dhidden, float* dweight, void* dweight_bf16,
                              float* dbias, b200trl_stream_t stream);

/* The same operator with the PADDING TRIMMED.  seq_rows_host: HOST memory, B entries; seq_rows_host[b] = number of
 * leading rows of sequence b that can carry a non-zero mask (index of its last unmasked token + 1; 0 for a fully masked
 * sequence).  Every sequence becomes its own chunk of seq_rows_host[b] rows, so the three contractions do
 * sum(seq_rows) / (B*T) of the dense work (a batch padded to T = 2048 with lengths in [T/2, T]: 0.75).  Loss, metrics,
 * dW, dbias as the dense call (order of fp32 sums aside); dhidden rows, logp and entropy behind seq_rows_host[b] are
 * zeros.  The caller needs the lengths on the host (one device->host read of B integers per call). */
int b200trl_fused_linear_grpo_trimmed(const void* hidden, const void* weight, const void* bias, int64_t B, int64_t T,
                                      int64_t H, int64_t V, const int64_t* ids, const int32_t* mask,
                                      const float* advantages, const float* old_logp, const float* ref_logp,
                                      const b200trl_grpo_cfg* cfg, float inv_temperature, const int64_t* seq_rows_host,
                                      void* workspace, float* logp, float* entropy, float* loss, float* metrics,
                                      void* dhidden, float* dweight, void* dweight_bf16, float* dbias,
                                      b200trl_stream_t stream);

/* ---- a-12: masked_mean / masked_var / masked_whiten (trl/core.py:43-76) ------------------------
 * stats fp32 [3] = {mean, unbiased var, count}; out (whitened, fp32 [n]) may be NULL.
 * workspace >= b200trl_masked_workspace_bytes(n), any contents. */
int64_t b200trl_masked_workspace_bytes(int64_t n);
int b200trl_masked_whiten(const float* values, const uint8_t* mask, int64_t n, int shift_mean, void* workspace,
                          float* out, float* stats, b200trl_stream_t stream);

/* ---- f-4: the integer masks either side of the loss (bit-exact index work, one launch each) ------------------
 * first_true_indices (trl/trainer/utils.py:877-897): out[r] = first t with bools[r, t] != 0, else T.
 * bools: uint8 [rows, T] (a torch.bool tensor's storage). */
int b200trl_first_true_indices(const uint8_t* bools, int64_t rows, int64_t T, int64_t* out, b200trl_stream_t stream);
/* "Mask everything after the first EOS token" (trl/trainer/grpo_trainer.py:1812-1817): eos_idx[b] (nullable) =
 * first t with completion_ids[b, t] == eos_token_id, else T; completion_mask[b, t] = (t <= eos_idx[b]) as int32. */
int b200trl_completion_mask(const int64_t* completion_ids, int64_t B, int64_t T, int64_t eos_token_id,
                            int32_t* completion_mask, int64_t* eos_idx, b200trl_stream_t stream);
/* truncate_response (utils.py:1036-1056) fused with the sequence length the PPO / RLOO loops derive from it
 * (ppo_trainer.py:455-464, rloo_trainer.py:347-355): postprocessed (nullable) = responses with everything after
 * the first stop token replaced by pad (has_stop_token = 0: unchanged, the `stop_token_id is None` branch);
 * sequence_length[b] (nullable) = first_true_indices(postprocessed[b] == pad) - 1. */
int b200trl_truncate_response(const int64_t* responses, int64_t B, int64_t T, int has_stop_token,
                              int64_t stop_token_id, int64_t pad_token_id, int64_t* postprocessed,
                              int64_t* sequence_length, b200trl_stream_t stream);

/* buf *= (*actual / expected) unless they are equal; used when autograd's grad_output differs from
 * the grad_scale assumed in the fused forward.  No host sync. */
int b200trl_rescale_if_needed(void* buf, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                              const float* actual, float expected, b200trl_stream_t stream);
/* The same for a two-level layout: row r = (b, t) at b * batch_stride + t * row_stride with rows_per_batch rows per
 * batch (a dlogits buffer that mirrors a strided view of the model output); rows_per_batch == 0: flat. */
int b200trl_rescale_if_needed_batched(void* buf, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                                      int64_t rows_per_batch, int64_t batch_stride, const float* actual, float expected,
                                      b200trl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200TRL_H_ */
