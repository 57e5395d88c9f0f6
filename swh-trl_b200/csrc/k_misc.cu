// a-4  b200trl_entropy_quantile_mask : get_high_entropy_mask, trl/trainer/grpo_trainer.py:341-364
//      exact torch.quantile(linear) of the non-pad entropies by radix select, then the >= compare.
//      b200trl_rescale_if_needed     : device-side fix-up when autograd's grad_output != assumed grad_scale.
#include "common.cuh"

namespace b200trl {
namespace {

constexpr int kSelBlock = 1024;

__device__ __forceinline__ uint32_t ordered_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// Single CTA: the data is [B,T] fp32 (<= a few MB, L2 resident); 4 histogram passes pick the k-th smallest key.
__global__ void __launch_bounds__(kSelBlock) entropy_quantile_kernel(const float* __restrict__ ent,
                                                                     const int32_t* __restrict__ mask, int64_t n,
                                                                     float q, uint8_t* __restrict__ out,
                                                                     float* __restrict__ thr_out) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_count;
    __shared__ unsigned int s_prefix, s_rank, s_any_nan, s_min_gt;
    __shared__ unsigned long long s_cnt_le;
    const int tid = threadIdx.x;

    // count non-pad entries, detect NaN
    if (tid == 0) {
        s_count = 0ull;
        s_any_nan = 0u;
    }
    __syncthreads();
    {
        unsigned long long c = 0;
        unsigned int nan = 0;
        for (int64_t i = tid; i < n; i += kSelBlock) {
            if (mask[i] != 0) {
                ++c;
                nan |= isnan(ent[i]) ? 1u : 0u;
            }
        }
        atomicAdd(&s_count, c);
        if (nan) atomicOr(&s_any_nan, 1u);
    }
    __syncthreads();
    const unsigned long long cnt = s_count;
    if (cnt == 0ull) {  // :358-359 — no non-pad token: all False
        for (int64_t i = tid; i < n; i += kSelBlock) out[i] = 0;
        if (thr_out && tid == 0) thr_out[0] = __int_as_float(0x7fc00000);
        return;
    }
    // torch.quantile: rank = q * (cnt - 1) evaluated in fp32, lerp between floor and ceil ranks
    const float rank = q * static_cast<float>(cnt - 1ull);
    const float rank_lo_f = floorf(rank);
    const unsigned long long k_lo = static_cast<unsigned long long>(rank_lo_f);
    const unsigned long long k_hi = static_cast<unsigned long long>(ceilf(rank));
    const float w = rank - rank_lo_f;

    // radix select of the k_lo-th smallest (0-based)
    if (tid == 0) {
        s_prefix = 0u;
        s_rank = static_cast<unsigned int>(k_lo);  // cnt < 2^32 for any [B,T] that fits here
    }
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t hi_mask = (shift == 24) ? 0u : (0xffffffffu << (shift + 8));
        for (int64_t i = tid; i < n; i += kSelBlock) {
            if (mask[i] != 0) {
                const uint32_t key = ordered_key(ent[i]);
                if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int r = s_rank, b = 0;
            for (; b < 256; ++b) {
                if (r < hist[b]) break;
                r -= hist[b];
            }
            s_rank = r;
            s_prefix = prefix | (b << shift);
        }
        __syncthreads();
    }
    const uint32_t key_lo = s_prefix;
    // the next order statistic: equal to key_lo if duplicated far enough, else the smallest key above it
    if (tid == 0) {
        s_cnt_le = 0ull;
        s_min_gt = 0xffffffffu;
    }
    __syncthreads();
    {
        unsigned long long le = 0;
        uint32_t mn = 0xffffffffu;
        for (int64_t i = tid; i < n; i += kSelBlock) {
            if (mask[i] != 0) {
                const uint32_t key = ordered_key(ent[i]);
                if (key <= key_lo) ++le;
                else mn = min(mn, key);
            }
        }
        atomicAdd(&s_cnt_le, le);
        atomicMin(&s_min_gt, mn);
    }
    __syncthreads();
    const float x_lo = key_to_float(key_lo);
    const float x_hi = (k_hi == k_lo || s_cnt_le > k_hi) ? x_lo : key_to_float(s_min_gt);
    // at::lerp: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
    const float diff = x_hi - x_lo;
    float thr = (w < 0.5f) ? fmaf(w, diff, x_lo) : x_hi - diff * (1.f - w);
    if (s_any_nan) thr = __int_as_float(0x7fc00000);
    if (thr_out && tid == 0) thr_out[0] = thr;
    // (entropies * mask) >= threshold, and not padding (:361-364)
    for (int64_t i = tid; i < n; i += kSelBlock) {
        const bool m = mask[i] != 0;
        out[i] = (m && (ent[i] * 1.0f >= thr)) ? 1 : 0;
    }
}

template <typename T>
__global__ void rescale_kernel(T* buf, int64_t n_rows, int64_t vocab, int64_t row_stride, const float* actual,
                               float expected) {
    const float a = actual[0];
    if (a == expected) return;  // the common case: nothing to do, the launch costs a few microseconds
    const float f = a / expected;
    const int64_t total = n_rows * vocab;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / vocab, c = i - r * vocab;
        T* p = buf + r * row_stride + c;
        ElemTraits<T>::store(p, ElemTraits<T>::load(p) * f);
    }
}

}  // namespace
}  // namespace b200trl

using namespace b200trl;

extern "C" int64_t b200trl_entropy_quantile_workspace_bytes(int64_t) { return 64; }

extern "C" int b200trl_entropy_quantile_mask(const float* entropies, const int32_t* mask, int64_t n, float threshold,
                                             void* /*workspace*/, uint8_t* out_mask, float* out_threshold,
                                             b200trl_stream_t stream) {
    B200TRL_REQUIRE(entropies && mask && out_mask, B200TRL_E_INVALID, "entropy_quantile_mask: null pointer");
    B200TRL_REQUIRE(n > 0 && n < (int64_t(1) << 32), B200TRL_E_INVALID, "entropy_quantile_mask: bad size %lld",
                    (long long)n);
    B200TRL_REQUIRE(threshold >= 0.f && threshold <= 1.f, B200TRL_E_INVALID,
                    "entropy_quantile_mask: quantile %f outside [0,1]", threshold);
    entropy_quantile_kernel<<<1, kSelBlock, 0, as_stream(stream)>>>(entropies, mask, n, threshold, out_mask,
                                                                    out_threshold);
    return check_launch("entropy_quantile_kernel");
}

extern "C" int b200trl_rescale_if_needed(void* buf, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                                         const float* actual, float expected, b200trl_stream_t stream) {
    B200TRL_REQUIRE(buf && actual, B200TRL_E_INVALID, "rescale_if_needed: null pointer");
    B200TRL_REQUIRE(expected != 0.f, B200TRL_E_INVALID, "rescale_if_needed: expected scale is zero");
    const int block = 256;
    const unsigned grid = static_cast<unsigned>(num_sms() * 8);
    cudaStream_t s = as_stream(stream);
    switch (dtype) {
        case B200TRL_BF16:
            rescale_kernel<<<grid, block, 0, s>>>(static_cast<__nv_bfloat16*>(buf), n_rows, vocab, row_stride, actual,
                                                  expected);
            break;
        case B200TRL_F16:
            rescale_kernel<<<grid, block, 0, s>>>(static_cast<__half*>(buf), n_rows, vocab, row_stride, actual, expected);
            break;
        case B200TRL_F32:
            rescale_kernel<<<grid, block, 0, s>>>(static_cast<float*>(buf), n_rows, vocab, row_stride, actual, expected);
            break;
        case B200TRL_F64:
            rescale_kernel<<<grid, block, 0, s>>>(static_cast<double*>(buf), n_rows, vocab, row_stride, actual, expected);
            break;
        default: set_error("rescale_if_needed: unknown dtype %d", dtype); return B200TRL_E_UNSUPPORTED;
    }
    return check_launch("rescale_kernel");
}
