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

}  // namespace
}  // namespace b200trl

using namespace b200trl;

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
