// RLOO (the next consumer of the same primitives):
//   b200trl_rloo_rewards_advantages  trl/trainer/rloo_trainer.py:397-441  KL penalty, (normalised) scores,
//                                    leave-one-out baseline over reshape(rloo_k, -1), (normalised) advantages
//   b200trl_rloo_loss                rloo_trainer.py:476-507  sequence ratio from summed log-probs, clipped
//                                    surrogate, stats, per-token d(loss)/d(logp) for the K1 backward
#include "token_math.cuh"

namespace b200trl {
namespace {

constexpr int kBlock = 256;

struct Ws {
    unsigned int counter;
    unsigned int pad[3];
    float rows[1];  // [mb][8]
};

__device__ __forceinline__ float block_sum1(float v, float* red) {
    float a[1] = {v};
    block_sum<1, 1024>(a, red);
    return a[0];
}

// one CTA of 1024 threads; B is at most a few thousand sequences
__global__ void __launch_bounds__(1024) rloo_adv_kernel(const float* __restrict__ lp, const float* __restrict__ rlp,
                                                        const float* __restrict__ scores,
                                                        const int64_t* __restrict__ seq_len, int64_t B, int64_t T,
                                                        float kl_coef, int64_t k, int norm_reward, float clip,
                                                        int norm_adv, int token_level, float* __restrict__ adv,
                                                        float* __restrict__ rlhf, float* __restrict__ non_score,
                                                        float* __restrict__ lp_f, float* __restrict__ rlp_f) {
    extern __shared__ float sh[];  // [B] rlhf reward
    __shared__ float red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // score normalisation (:407-409): unbiased std over the batch
    float s_mean = 0.f, s_inv = 1.f;
    if (norm_reward) {
        float s = 0.f;
        for (int64_t b = tid; b < B; b += 1024) s += scores[b];
        s_mean = block_sum1(s, red) / static_cast<float>(B);
        float q = 0.f;
        for (int64_t b = tid; b < B; b += 1024) {
            const float d = scores[b] - s_mean;
            q += d * d;
        }
        const float var = block_sum1(q, red) / static_cast<float>(B - 1);
        s_inv = 1.f / (sqrtf(var) + 1e-8f);
    }
    // per-sequence KL sums, one warp per row (:404, 412-431)
    for (int64_t b = warp; b < B; b += 32) {
        const int64_t len = seq_len[b];
        float kl_sum = 0.f;
        for (int64_t t = lane; t < T; t += 32) {
            const bool pad = t > len;
            const float a = pad ? 1.0f : lp[b * T + t];  // INVALID_LOGPROB (:399-400)
            const float r = pad ? 1.0f : rlp[b * T + t];
            if (lp_f) lp_f[b * T + t] = a;
            if (rlp_f) rlp_f[b * T + t] = r;
            kl_sum += token_level ? -kl_coef * (a - r) : (a - r);
        }
        kl_sum = warp_sum(kl_sum);
        if (lane == 0) {
            float sc = scores[b];
            if (norm_reward) sc = fminf(fmaxf((sc - s_mean) * s_inv, -clip), clip);
            const float ns = token_level ? kl_sum : -kl_coef * kl_sum;
            non_score[b] = ns;
            sh[b] = ns + sc;
            rlhf[b] = ns + sc;
        }
    }
    __syncthreads();
    // leave-one-out baseline over the rloo_k samples of a prompt: sample i of prompt p is row i * (B/k) + p (:434-437)
    const int64_t P = B / k;
    float a_sum = 0.f;
    for (int64_t b = tid; b < B; b += 1024) {
        const int64_t p = b % P;
        float col = 0.f;
        for (int64_t i = 0; i < k; ++i) col += sh[i * P + p];
        const float r = sh[b];
        const float a = r - (col - r) / static_cast<float>(k - 1);
        adv[b] = a;
        a_sum += a;
    }
    if (norm_adv) {  // :440-441
        const float mean = block_sum1(a_sum, red) / static_cast<float>(B);
        __syncthreads();
        float q = 0.f;
        for (int64_t b = tid; b < B; b += 1024) {
            const float d = adv[b] - mean;
            q += d * d;
        }
        const float var = block_sum1(q, red) / static_cast<float>(B - 1);
        const float inv = 1.f / (sqrtf(var) + 1e-8f);
        for (int64_t b = tid; b < B; b += 1024) adv[b] = (adv[b] - mean) * inv;
    }
}

struct LossArgs {
    const float* new_lp;
    const float* old_lp;
    const float* adv;
    const float* entropy;
    const int64_t* seq_len;
    int64_t mb, T;
    float clip_lo, clip_hi, grad_scale;
    Ws* ws;
    float* stats;
    float* g;
};

__global__ void __launch_bounds__(kBlock) rloo_loss_kernel(const LossArgs a) {
    __shared__ float red[4 * 32];
    __shared__ bool is_last;
    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t len = a.seq_len[b];
    float v[4] = {0.f, 0.f, 0.f, 0.f};  // sum new, sum old, sum exp(new-old), sum entropy
    for (int64_t t = tid; t < a.T; t += kBlock) {
        const int64_t i = b * a.T + t;
        const float n = (t > len) ? 1.0f : a.new_lp[i];  // :471-473
        const float o = a.old_lp[i];
        v[0] += n;
        v[1] += o;
        v[2] += expf(n - o);  // new_ratio (:476)
        if (a.entropy) v[3] += a.entropy[i];
    }
    block_sum<4, kBlock>(v, red);
    float pg, dpg, clipped, ratio, diff;
    ppo_policy(v[0], v[1], a.adv[b], a.clip_lo, a.clip_hi, pg, dpg, clipped, ratio, diff);  // :477-485 on the sums
    if (a.g) {
        const float gb = dpg / static_cast<float>(a.mb) * a.grad_scale;  // pg_loss_max.mean() (:486)
        for (int64_t t = tid; t < a.T; t += kBlock) a.g[b * a.T + t] = (t > len) ? 0.f : gb;
    }
    if (tid == 0) {
        float* r = a.ws->rows + b * 8;
        r[0] = pg;
        r[1] = clipped;
        r[2] = diff * diff;
        r[3] = v[3];
        r[4] = v[2];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int prev = atomicAdd(&a.ws->counter, 1u);
        is_last = (prev == gridDim.x - 1);
        if (is_last) a.ws->counter = 0u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid < 32) {
        double acc[5] = {0, 0, 0, 0, 0};
        for (int64_t r = tid; r < a.mb; r += 32) {
            const volatile float* p = a.ws->rows + r * 8;
#pragma unroll
            for (int k = 0; k < 5; ++k) acc[k] += static_cast<double>(p[k]);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[k] = warp_sum(acc[k]);
        if (tid == 0) {
            const double n = static_cast<double>(a.mb), nt = n * static_cast<double>(a.T);
            a.stats[B200TRL_R_LOSS] = static_cast<float>(acc[0] / n);
            a.stats[B200TRL_R_PG_CLIPFRAC] = static_cast<float>(acc[1] / n);
            a.stats[B200TRL_R_APPROXKL] = static_cast<float>(0.5 * acc[2] / n);
            a.stats[B200TRL_R_ENTROPY] = static_cast<float>(acc[3] / nt);
            a.stats[B200TRL_R_RATIO] = static_cast<float>(acc[4] / nt);
            a.stats[5] = a.stats[6] = a.stats[7] = 0.f;
        }
    }
}

}  // namespace
}  // namespace b200trl

using namespace b200trl;

extern "C" int b200trl_rloo_rewards_advantages(const float* logprobs, const float* ref_logprobs, const float* scores,
                                               const int64_t* sequence_lengths, int64_t B, int64_t T, float kl_coef,
                                               int64_t rloo_k, int normalize_reward, float reward_clip_range,
                                               int normalize_advantage, int token_level_kl, float* advantages,
                                               float* rlhf_reward, float* non_score_reward, float* logprobs_f,
                                               float* ref_logprobs_f, b200trl_stream_t stream) {
    B200TRL_REQUIRE(logprobs && ref_logprobs && scores && sequence_lengths && advantages && rlhf_reward && non_score_reward,
                    B200TRL_E_INVALID, "rloo_rewards_advantages: null pointer");
    B200TRL_REQUIRE(B > 0 && T > 0 && rloo_k >= 2 && B % rloo_k == 0, B200TRL_E_INVALID,
                    "rloo_rewards_advantages: batch %lld must be a positive multiple of rloo_k %lld >= 2", (long long)B,
                    (long long)rloo_k);
    B200TRL_REQUIRE(B <= 12000, B200TRL_E_UNSUPPORTED, "rloo_rewards_advantages: batch %lld exceeds the single-CTA limit",
                    (long long)B);
    rloo_adv_kernel<<<1, 1024, static_cast<size_t>(B) * sizeof(float), as_stream(stream)>>>(
        logprobs, ref_logprobs, scores, sequence_lengths, B, T, kl_coef, rloo_k, normalize_reward, reward_clip_range,
        normalize_advantage, token_level_kl, advantages, rlhf_reward, non_score_reward, logprobs_f, ref_logprobs_f);
    return check_launch("rloo_adv_kernel");
}

extern "C" int b200trl_rloo_loss(const float* new_logprobs, const float* old_logprobs, const float* advantages,
                                 const float* entropy, const int64_t* sequence_lengths, int64_t mb, int64_t T,
                                 float cliprange, float grad_scale, void* workspace, float* stats, float* g,
                                 b200trl_stream_t stream) {
    B200TRL_REQUIRE(new_logprobs && old_logprobs && advantages && sequence_lengths && workspace && stats,
                    B200TRL_E_INVALID, "rloo_loss: null pointer");
    B200TRL_REQUIRE(mb > 0 && T > 0 && mb <= 0x7fffffff, B200TRL_E_INVALID, "rloo_loss: bad shape");
    LossArgs a{new_logprobs, old_logprobs, advantages, entropy, sequence_lengths, mb, T,
               static_cast<float>(1.0 - static_cast<double>(cliprange)),
               static_cast<float>(1.0 + static_cast<double>(cliprange)), grad_scale, static_cast<Ws*>(workspace), stats, g};
    rloo_loss_kernel<<<static_cast<unsigned>(mb), kBlock, 0, as_stream(stream)>>>(a);
    return check_launch("rloo_loss_kernel");
}
