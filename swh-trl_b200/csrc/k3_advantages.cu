// K3: group-relative advantages — trl/trainer/grpo_trainer.py:1917-1938.
// Input is the all-gathered, rank-major rewards_per_func [B_global, n_funcs]; a group is G consecutive rows of it
// (:1921 `view(-1, G)`), whichever ranks they came from.  One thread per group; B_global is a few hundred rows.
#include "common.cuh"

namespace b200trl {
namespace {

__global__ void group_advantages_kernel(const float* __restrict__ rpf, const float* __restrict__ w, int64_t n_groups,
                                        int64_t F, int64_t G, int scale, int64_t local_offset, int64_t local_count,
                                        float* __restrict__ rewards, float* __restrict__ adv_all,
                                        float* __restrict__ adv_local, float* __restrict__ mean_out,
                                        float* __restrict__ std_out, uint8_t* __restrict__ zero_out) {
    const int64_t grp = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (grp >= n_groups) return;
    const int64_t r0 = grp * G;
    // rewards = nansum(rewards_per_func * weights, dim=1)   (:1918)
    double sum = 0.0;
    for (int64_t i = 0; i < G; ++i) {
        float acc = 0.f;
        for (int64_t f = 0; f < F; ++f) {
            const float v = rpf[(r0 + i) * F + f] * w[f];
            if (!isnan(v)) acc += v;
        }
        rewards[r0 + i] = acc;
        sum += static_cast<double>(acc);
    }
    const float mean = static_cast<float>(sum / static_cast<double>(G));  // :1921
    double ss = 0.0;
    for (int64_t i = 0; i < G; ++i) {
        const double d = static_cast<double>(rewards[r0 + i]) - static_cast<double>(mean);
        ss += d * d;
    }
    // unbiased std (:1922); G == 1 gives 0/0 = NaN exactly as torch.std does
    const float sd = static_cast<float>(sqrt(ss / static_cast<double>(G - 1)));
    mean_out[grp] = mean;
    std_out[grp] = sd;
    zero_out[grp] = (fabsf(sd) <= 1e-8f) ? 1 : 0;  // torch.isclose(std, 0): atol 1e-8 (:1923)
    for (int64_t i = 0; i < G; ++i) {
        float a = rewards[r0 + i] - mean;  // :1928
        if (scale) a = a / (sd + 1e-4f);   // :1930
        adv_all[r0 + i] = a;
        const int64_t l = r0 + i - local_offset;  // process_slice (:1933-1938)
        if (adv_local && l >= 0 && l < local_count) adv_local[l] = a;
    }
}

// ---- f-4: the logging block of _generate_and_score_completions (grpo_trainer.py:1942-1972) in one launch -------------
// packed: the all-gathered int64 vector of every rank, [world][1 + 2 * B_local] = {sum(attention_mask),
// completion_lengths[B_local], terminated_with_eos[B_local]}.  One CTA; sums in double (the lengths are integers: exact).
constexpr int kGenBlock = 256;
constexpr int kGenFixed = 12;  // out[0..11], then (mean, std) per reward function

__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = (threadIdx.x < kGenBlock / 32) ? red[threadIdx.x] : 0.0;
    t = warp_sum(t);
    t = __shfl_sync(0xffffffffu, t, 0);
    __syncthreads();
    if (threadIdx.x == 0) red[0] = t;
    __syncthreads();
    t = red[0];
    __syncthreads();
    return t;
}
__device__ __forceinline__ double block_min_d(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = red[0];
    for (int w = 1; w < kGenBlock / 32; ++w) t = fmin(t, red[w]);
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kGenBlock) generation_stats_kernel(
    const int64_t* __restrict__ packed, int64_t world, int64_t b_local, const float* __restrict__ rpf, int64_t n_funcs,
    const float* __restrict__ mean_g, const float* __restrict__ std_g, const uint8_t* __restrict__ zero_g, int64_t n_groups,
    double* __restrict__ out) {
    __shared__ double red[kGenBlock / 32];
    const int tid = threadIdx.x;
    const int64_t stride = 1 + 2 * b_local, Bg = world * b_local;
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    double tok = 0.0;
    for (int64_t r = tid; r < world; r += kGenBlock) tok += static_cast<double>(packed[r * stride]);
    double s = 0.0, mn = kInf, mx = -kInf, ts = 0.0, tn = 0.0, tmn = kInf, tmx = -kInf;
    for (int64_t i = tid; i < Bg; i += kGenBlock) {
        const int64_t r = i / b_local, j = i - r * b_local;
        const double len = static_cast<double>(packed[r * stride + 1 + j]);
        const bool term = packed[r * stride + 1 + b_local + j] != 0;
        s += len;
        mn = fmin(mn, len);
        mx = fmax(mx, len);
        if (term) {
            ts += len;
            tn += 1.0;
            tmn = fmin(tmn, len);
            tmx = fmax(tmx, len);
        }
    }
    tok = block_sum_d(tok, red);
    s = block_sum_d(s, red);
    ts = block_sum_d(ts, red);
    tn = block_sum_d(tn, red);
    mn = block_min_d(mn, red);
    mx = -block_min_d(-mx, red);
    tmn = block_min_d(tmn, red);
    tmx = -block_min_d(-tmx, red);
    double gm = 0.0, gs = 0.0, gz = 0.0;
    for (int64_t g = tid; g < n_groups; g += kGenBlock) {
        gm += static_cast<double>(mean_g[g]);
        gs += static_cast<double>(std_g[g]);
        gz += zero_g[g] ? 1.0 : 0.0;
    }
    gm = block_sum_d(gm, red);
    gs = block_sum_d(gs, red);
    gz = block_sum_d(gz, red);
    if (tid == 0) {
        const double n = static_cast<double>(Bg);
        out[0] = tok;                                  // :1943 (summed over ranks)
        out[1] = s / n;                                // completions/mean_length (:1947)
        out[2] = mn;                                   // :1948
        out[3] = mx;                                   // :1949
        out[4] = 1.0 - tn / n;                         // completions/clipped_ratio (:1954)
        out[5] = tn > 0.0 ? ts / tn : 0.0;             // :1956-1958: zeros(1) when nothing terminated
        out[6] = tn > 0.0 ? tmn : 0.0;
        out[7] = tn > 0.0 ? tmx : 0.0;
        out[8] = gm / static_cast<double>(n_groups);   // reward (:1966; the mean of the repeated vector is the group mean)
        out[9] = gs / static_cast<double>(n_groups);   // reward_std (:1967)
        out[10] = gz / static_cast<double>(n_groups);  // frac_reward_zero_std (:1968)
        out[11] = 0.0;
    }
    // per reward function: nanmean and nanstd (grpo_trainer.py:196-211) over the samples the function applied to
    for (int64_t f = 0; f < n_funcs; ++f) {
        double fs = 0.0, fc = 0.0;
        for (int64_t i = tid; i < Bg; i += kGenBlock) {
            const float v = rpf[i * n_funcs + f];
            if (!isnan(v)) {
                fs += static_cast<double>(v);
                fc += 1.0;
            }
        }
        fs = block_sum_d(fs, red);
        fc = block_sum_d(fc, red);
        const double m = fs / fc;  // 0 / 0 = NaN, as torch.nanmean of an all-NaN column
        double q = 0.0;
        for (int64_t i = tid; i < Bg; i += kGenBlock) {
            const float v = rpf[i * n_funcs + f];
            if (!isnan(v)) {
                const double d = static_cast<double>(v) - m;
                q += d * d;
            }
        }
        q = block_sum_d(q, red);
        if (tid == 0) {
            out[kGenFixed + 2 * f] = m;
            out[kGenFixed + 2 * f + 1] = sqrt((q / fc) * (fc / (fc - 1.0)));  // Bessel as written at :210 (count 1 -> NaN)
        }
    }
}

}  // namespace
}  // namespace b200trl

using namespace b200trl;

extern "C" int b200trl_generation_stats(const int64_t* packed, int64_t world, int64_t b_local,
                                        const float* rewards_per_func, int64_t n_funcs, const float* mean_grouped,
                                        const float* std_grouped, const uint8_t* is_std_zero, int64_t n_groups,
                                        double* out, b200trl_stream_t stream) {
    B200TRL_REQUIRE(packed && rewards_per_func && mean_grouped && std_grouped && is_std_zero && out, B200TRL_E_INVALID,
                    "generation_stats: null pointer");
    B200TRL_REQUIRE(world > 0 && b_local > 0 && n_funcs > 0 && n_groups > 0, B200TRL_E_INVALID,
                    "generation_stats: bad sizes");
    generation_stats_kernel<<<1, kGenBlock, 0, as_stream(stream)>>>(packed, world, b_local, rewards_per_func, n_funcs,
                                                                   mean_grouped, std_grouped, is_std_zero, n_groups, out);
    return check_launch("generation_stats_kernel");
}

extern "C" int b200trl_group_advantages(const float* rewards_per_func, const float* weights, int64_t B_global,
                                        int64_t n_funcs, int64_t G, int scale_rewards, int64_t local_offset,
                                        int64_t local_count, float* rewards, float* adv_all, float* adv_local,
                                        float* mean, float* std, uint8_t* is_std_zero, b200trl_stream_t stream) {
    B200TRL_REQUIRE(rewards_per_func && weights && rewards && adv_all && mean && std && is_std_zero, B200TRL_E_INVALID,
                    "group_advantages: null pointer");
    B200TRL_REQUIRE(B_global > 0 && n_funcs > 0 && G > 0, B200TRL_E_INVALID, "group_advantages: bad sizes");
    B200TRL_REQUIRE(B_global % G == 0, B200TRL_E_INVALID,
                    "group_advantages: global batch %lld is not a multiple of num_generations %lld", (long long)B_global,
                    (long long)G);
    B200TRL_REQUIRE(local_offset >= 0 && local_count >= 0 && local_offset + local_count <= B_global, B200TRL_E_INVALID,
                    "group_advantages: local slice [%lld, +%lld) outside the global batch", (long long)local_offset,
                    (long long)local_count);
    const int64_t n_groups = B_global / G;
    const int block = 128;
    const unsigned grid = static_cast<unsigned>((n_groups + block - 1) / block);
    group_advantages_kernel<<<grid, block, 0, as_stream(stream)>>>(rewards_per_func, weights, n_groups, n_funcs, G,
                                                                  scale_rewards, local_offset, local_count, rewards,
                                                                  adv_all, adv_local, mean, std, is_std_zero);
    return check_launch("group_advantages_kernel");
}
