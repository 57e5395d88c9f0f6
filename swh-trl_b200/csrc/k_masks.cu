// f-4  The integer mask constructions on either side of the loss path: bit-exact index work.
//   b200trl_first_true_indices : first_true_indices, trl/trainer/utils.py:877-897
//   b200trl_completion_mask    : "mask everything after the first EOS", trl/trainer/grpo_trainer.py:1812-1817
//   b200trl_truncate_response  : truncate_response, utils.py:1036-1056 + sequence_length, ppo_trainer.py:464 /
//                                rloo_trainer.py:355
// One warp per row: each lane scans a strided share with an early exit, a shuffle-min picks the first hit.  The
// reference spends 5-8 [B,T] elementwise kernels (compare, int cast, argmax / min, any, arange, expand, compare,
// masked_fill) on each of these.
#include "common.cuh"

namespace b200trl {
namespace {

constexpr int kRowsPerBlock = 8;  // 8 warps per CTA, one row each

__device__ __forceinline__ int64_t warp_min(int64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// index of the first t with pred(t), T if none; warp-cooperative, all lanes get the result
template <typename Pred>
__device__ __forceinline__ int64_t first_hit(int64_t T, int lane, Pred pred) {
    int64_t first = T;
    for (int64_t t0 = 0; t0 < T; t0 += 32) {  // uniform trip count: the ballot needs every lane
        const int64_t t = t0 + lane;
        const bool hit = (t < T) && pred(t);
        const unsigned int b = __ballot_sync(0xffffffffu, hit);
        if (b) {
            first = t0 + (__ffs(b) - 1);
            break;
        }
    }
    return first;
}

__global__ void __launch_bounds__(kRowsPerBlock * 32) first_true_kernel(const uint8_t* __restrict__ bools, int64_t rows,
                                                                        int64_t T, int64_t* __restrict__ out) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const uint8_t* row = bools + r * T;
    const int64_t f = first_hit(T, lane, [&](int64_t t) { return row[t] != 0; });
    if (lane == 0) out[r] = f;
}

__global__ void __launch_bounds__(kRowsPerBlock * 32) completion_mask_kernel(const int64_t* __restrict__ ids,
                                                                             int64_t rows, int64_t T, int64_t eos,
                                                                             int32_t* __restrict__ mask,
                                                                             int64_t* __restrict__ eos_idx) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int64_t* row = ids + r * T;
    const int64_t f = first_hit(T, lane, [&](int64_t t) { return row[t] == eos; });  // T when the row has no EOS (:1814)
    if (lane == 0 && eos_idx) eos_idx[r] = f;
    for (int64_t t = lane; t < T; t += 32) mask[r * T + t] = (t <= f) ? 1 : 0;  // :1817, up to and including the EOS
}

__global__ void __launch_bounds__(kRowsPerBlock * 32) truncate_response_kernel(
    const int64_t* __restrict__ responses, int64_t rows, int64_t T, int has_stop, int64_t stop, int64_t pad,
    int64_t* __restrict__ out, int64_t* __restrict__ sequence_length) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * kRowsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int64_t* row = responses + r * T;
    // utils.py:1052-1055: everything AFTER the first stop token becomes pad (the stop token itself stays)
    const int64_t trunc = has_stop ? first_hit(T, lane, [&](int64_t t) { return row[t] == stop; }) : T;
    // ppo_trainer.py:464: first pad of the post-processed response, minus one
    const int64_t first_pad = first_hit(T, lane, [&](int64_t t) { return (t > trunc) || row[t] == pad; });
    if (out)
        for (int64_t t = lane; t < T; t += 32) out[r * T + t] = (t > trunc) ? pad : row[t];
    if (lane == 0 && sequence_length) sequence_length[r] = first_pad - 1;
}

int grid_for(int64_t rows) { return static_cast<int>((rows + kRowsPerBlock - 1) / kRowsPerBlock); }

}  // namespace
}  // namespace b200trl

using namespace b200trl;

extern "C" int b200trl_first_true_indices(const uint8_t* bools, int64_t rows, int64_t T, int64_t* out,
                                          b200trl_stream_t stream) {
    B200TRL_REQUIRE(rows >= 0 && T >= 0, B200TRL_E_INVALID, "first_true_indices: bad shape");
    if (rows == 0) return B200TRL_OK;
    B200TRL_REQUIRE(out && (bools || T == 0), B200TRL_E_INVALID, "first_true_indices: null pointer");
    first_true_kernel<<<grid_for(rows), kRowsPerBlock * 32, 0, as_stream(stream)>>>(bools, rows, T, out);
    return check_launch("first_true_kernel");
}

extern "C" int b200trl_completion_mask(const int64_t* completion_ids, int64_t B, int64_t T, int64_t eos_token_id,
                                       int32_t* completion_mask, int64_t* eos_idx, b200trl_stream_t stream) {
    B200TRL_REQUIRE(B >= 0 && T >= 0, B200TRL_E_INVALID, "completion_mask: bad shape");
    if (B == 0) return B200TRL_OK;
    B200TRL_REQUIRE((completion_ids && completion_mask) || T == 0, B200TRL_E_INVALID, "completion_mask: null pointer");
    completion_mask_kernel<<<grid_for(B), kRowsPerBlock * 32, 0, as_stream(stream)>>>(completion_ids, B, T, eos_token_id,
                                                                                      completion_mask, eos_idx);
    return check_launch("completion_mask_kernel");
}

extern "C" int b200trl_truncate_response(const int64_t* responses, int64_t B, int64_t T, int has_stop_token,
                                         int64_t stop_token_id, int64_t pad_token_id, int64_t* postprocessed,
                                         int64_t* sequence_length, b200trl_stream_t stream) {
    B200TRL_REQUIRE(B >= 0 && T >= 0, B200TRL_E_INVALID, "truncate_response: bad shape");
    if (B == 0) return B200TRL_OK;
    B200TRL_REQUIRE(responses || T == 0, B200TRL_E_INVALID, "truncate_response: null pointer");
    B200TRL_REQUIRE(postprocessed || sequence_length, B200TRL_E_INVALID, "truncate_response: nothing to do");
    truncate_response_kernel<<<grid_for(B), kRowsPerBlock * 32, 0, as_stream(stream)>>>(
        responses, B, T, has_stop_token, stop_token_id, pad_token_id, postprocessed, sequence_length);
    return check_launch("truncate_response_kernel");
}
