// a-4  b200trl_entropy_quantile_mask : get_high_entropy_mask, trl/trainer/grpo_trainer.py:341-364
//      exact torch.quantile(linear) of the non-pad entropies by radix select, then the >= compare.
//      Up to 131 072 tokens (the local batch of every BASELINE config) run in ONE 8-CTA thread-block cluster with the
//      keys held in registers: each token is read once, the 256-bin histograms of the four passes are merged in the
//      leader CTA's shared memory through DSMEM atomics and six cluster barriers replace the seven grid.sync() rounds
//      and the six passes over L2 of the cooperative kernel (75.8 us at 16 384 tokens); larger inputs use that one.
//      b200trl_rescale_if_needed     : device-side fix-up when autograd's grad_output != assumed grad_scale.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

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

// Cooperative multi-CTA radix select: every CTA histograms its share of the (L2-resident) [B,T] data into shared
// memory with warp-aggregated atomics (entropies cluster in a few exponent bins, plain atomics would serialise),
// merges into a global 256-bin histogram, and after a grid.sync every CTA walks the same histogram to the same
// bin.  4 passes pick the k-th smallest key exactly; one more pass finds the next order statistic.
struct QuantileWs {
    unsigned int hist[4][256];
    unsigned long long count;
    unsigned long long cnt_le;
    unsigned int any_nan;
    unsigned int min_gt;
};

__device__ __forceinline__ void hist_add(unsigned int* hist, uint32_t bin, bool active) {
    // warp-aggregated: one atomic per distinct bin per warp
    const unsigned int act = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned int peers = __match_any_sync(act, bin);
    const int leader = __ffs(peers) - 1;
    if ((threadIdx.x & 31) == leader) atomicAdd(&hist[bin], __popc(peers));
}

__global__ void __launch_bounds__(kSelBlock) entropy_quantile_kernel(const float* __restrict__ ent,
                                                                     const int32_t* __restrict__ mask, int64_t n,
                                                                     float q, QuantileWs* ws,
                                                                     uint8_t* __restrict__ out,
                                                                     float* __restrict__ thr_out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_rank;
    __shared__ unsigned int s_wtot[8];
    const int tid = threadIdx.x;
    const int64_t start = static_cast<int64_t>(blockIdx.x) * kSelBlock + tid;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kSelBlock;
    const int64_t n_round = ((n + stride - 1) / stride) * stride;  // uniform trip count: warp collectives inside

    // ---- pass 0: zero the workspace (block 0), count non-pad entries, detect NaN
    if (blockIdx.x == 0) {
        for (int i = tid; i < 4 * 256; i += kSelBlock) (&ws->hist[0][0])[i] = 0u;
        if (tid == 0) {
            ws->count = 0ull;
            ws->cnt_le = 0ull;
            ws->any_nan = 0u;
            ws->min_gt = 0xffffffffu;
        }
    }
    __threadfence();
    grid.sync();
    {
        unsigned long long c = 0;
        unsigned int nan = 0;
        for (int64_t i = start; i < n; i += stride) {
            if (mask[i] != 0) {
                ++c;
                nan |= isnan(ent[i]) ? 1u : 0u;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, o);
            nan |= __shfl_xor_sync(0xffffffffu, nan, o);
        }
        if ((tid & 31) == 0) {
            if (c) atomicAdd(&ws->count, c);
            if (nan) atomicOr(&ws->any_nan, 1u);
        }
    }
    __threadfence();
    grid.sync();
    const unsigned long long cnt = *reinterpret_cast<volatile unsigned long long*>(&ws->count);
    if (cnt == 0ull) {  // :358-359 — no non-pad token: all False
        for (int64_t i = start; i < n; i += stride) out[i] = 0;
        if (thr_out && blockIdx.x == 0 && tid == 0) thr_out[0] = __int_as_float(0x7fc00000);
        return;
    }
    // torch.quantile: rank = q * (cnt - 1) evaluated in fp32, lerp between floor and ceil ranks
    const float rank = q * static_cast<float>(cnt - 1ull);
    const float rank_lo_f = floorf(rank);
    const unsigned long long k_lo = static_cast<unsigned long long>(rank_lo_f);
    const unsigned long long k_hi = static_cast<unsigned long long>(ceilf(rank));
    const float w = rank - rank_lo_f;

    // ---- radix select of the k_lo-th smallest (0-based), 8 bits per pass
    uint32_t prefix = 0u;
    unsigned int r = static_cast<unsigned int>(k_lo);
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        const uint32_t hi_mask = (pass == 0) ? 0u : (0xffffffffu << (shift + 8));
        for (int64_t i = start; i < n_round; i += stride) {
            bool active = false;
            uint32_t bin = 0;
            if (i < n && mask[i] != 0) {
                const uint32_t key = ordered_key(ent[i]);
                active = ((key & hi_mask) == prefix);
                bin = (key >> shift) & 0xffu;
            }
            hist_add(hist, bin, active);
        }
        __syncthreads();
        if (tid < 256 && hist[tid]) atomicAdd(&ws->hist[pass][tid], hist[tid]);
        __threadfence();
        grid.sync();
        // every CTA finds the same bin in the same global histogram: 256 threads fetch one bin each (a serial walk
        // by one thread costs up to 256 dependent L2 round trips per pass), a two-level shuffle scan gives every
        // bin its exclusive prefix, and the one bin with prefix <= r < prefix + count publishes the new state
        if (tid < 256) {
            const unsigned int h = reinterpret_cast<const volatile unsigned int*>(ws->hist[pass])[tid];
            unsigned int inc = h;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
                if ((tid & 31) >= o) inc += up;
            }
            if ((tid & 31) == 31) s_wtot[tid >> 5] = inc;
            // (only warps 0..7 are in here, so a named barrier over these 256 threads orders the two levels)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            unsigned int base = 0;
            for (int wv = 0; wv < (tid >> 5); ++wv) base += s_wtot[wv];
            const unsigned int excl = base + inc - h;
            if (h != 0u && r >= excl && r < excl + h) {
                s_rank = r - excl;
                s_prefix = prefix | (static_cast<uint32_t>(tid) << shift);
            }
        }
        __syncthreads();
        r = s_rank;
        prefix = s_prefix;
        __syncthreads();
    }
    const uint32_t key_lo = prefix;
    // ---- the next order statistic: equal to key_lo if duplicated far enough, else the smallest key above it
    {
        unsigned long long le = 0;
        uint32_t mn = 0xffffffffu;
        for (int64_t i = start; i < n; i += stride) {
            if (mask[i] != 0) {
                const uint32_t key = ordered_key(ent[i]);
                if (key <= key_lo) ++le;
                else mn = min(mn, key);
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            le += __shfl_xor_sync(0xffffffffu, le, o);
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if ((tid & 31) == 0) {
            if (le) atomicAdd(&ws->cnt_le, le);
            atomicMin(&ws->min_gt, mn);
        }
    }
    __threadfence();
    grid.sync();
    const unsigned long long cnt_le = *reinterpret_cast<volatile unsigned long long*>(&ws->cnt_le);
    const uint32_t min_gt = *reinterpret_cast<volatile unsigned int*>(&ws->min_gt);
    const float x_lo = key_to_float(key_lo);
    const float x_hi = (k_hi == k_lo || cnt_le > k_hi) ? x_lo : key_to_float(min_gt);
    // at::lerp: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
    const float diff = x_hi - x_lo;
    float thr = (w < 0.5f) ? fmaf(w, diff, x_lo) : x_hi - diff * (1.f - w);
    if (*reinterpret_cast<volatile unsigned int*>(&ws->any_nan)) thr = __int_as_float(0x7fc00000);
    if (thr_out && blockIdx.x == 0 && tid == 0) thr_out[0] = thr;
    // (entropies * mask) >= threshold, and not padding (:361-364)
    for (int64_t i = start; i < n; i += stride) {
        const bool m = mask[i] != 0;
        out[i] = (m && (ent[i] * 1.0f >= thr)) ? 1 : 0;
    }
}


// ------------------------------------------------------------------ single-cluster variant (n <= 131 072)
constexpr int kQCtas = 8;
constexpr int kQPer = 16;  // keys per thread: 8 CTAs x 1024 threads x 16 = 131 072 tokens

// sum (a) and or / min (b) over the cluster: every CTA posts into every CTA's row, one cluster barrier
__device__ __forceinline__ void cluster_xchg(cg::cluster_group& cluster, unsigned long long (*row)[2], unsigned long long a,
                                             unsigned long long b, unsigned long long* s_red) {
    // block level: warp shuffles, then warp 0 over the 32 warp partials
    const int tid = threadIdx.x;
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((tid & 31) == 0) {
        s_red[2 * (tid >> 5)] = a;
        s_red[2 * (tid >> 5) + 1] = b;
    }
    __syncthreads();
    if (tid < 32) {
        a = s_red[2 * tid];
        b = s_red[2 * tid + 1];
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if (tid < kQCtas) {
            unsigned long long* dst = cluster.map_shared_rank(&row[cluster.block_rank()][0], tid);
            dst[0] = a;
            dst[1] = b;
        }
    }
    cluster.sync();
}

__global__ void __cluster_dims__(kQCtas, 1, 1) __launch_bounds__(kSelBlock, 1)
    entropy_quantile_cluster_kernel(const float* __restrict__ ent, const int32_t* __restrict__ mask, const int n,
                                    const float q, uint8_t* __restrict__ out, float* __restrict__ thr_out) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ unsigned int hist[256];
    __shared__ unsigned int ghist[4][256];  // the leader CTA's copy is the cluster-wide histogram of each pass
    __shared__ unsigned int s_prefix, s_rank;
    __shared__ unsigned int s_wtot[8];
    __shared__ unsigned long long s_red[64];
    __shared__ unsigned long long xq[2][kQCtas][2];
    const int tid = threadIdx.x;
    const int g = static_cast<int>(cluster.block_rank()) * kSelBlock + tid;  // thread index in the cluster
    constexpr int kStride = kQCtas * kSelBlock;
    const int per = (n + kStride - 1) / kStride;  // keys per thread, uniform over the cluster (<= kQPer)

    // ---- the only read of the inputs: keys of the non-pad tokens into registers
    uint32_t key[kQPer];
    unsigned int valid = 0u;
    unsigned long long c = 0ull, nan = 0ull;
#pragma unroll
    for (int j = 0; j < kQPer; ++j) {
        key[j] = 0u;
        const int i = j * kStride + g;
        if (j < per && i < n && mask[i] != 0) {
            const float e = ent[i];
            key[j] = ordered_key(e);
            valid |= 1u << j;
            ++c;
            nan |= isnan(e) ? 1ull : 0ull;
        }
    }
    for (int i = tid; i < 4 * 256; i += kSelBlock) (&ghist[0][0])[i] = 0u;
    // count and NaN flag (min of ~flag: 0 if any NaN)
    cluster_xchg(cluster, xq[0], c, nan ? 0ull : 1ull, s_red);  // also orders the zeroed ghist before remote atomics
    unsigned long long cnt = 0ull, no_nan = 1ull;
#pragma unroll
    for (int r = 0; r < kQCtas; ++r) {
        cnt += xq[0][r][0];
        no_nan = min(no_nan, xq[0][r][1]);
    }
    if (cnt == 0ull) {  // :358-359 — no non-pad token: all False
#pragma unroll
        for (int j = 0; j < kQPer; ++j) {
            const int i = j * kStride + g;
            if (j < per && i < n) out[i] = 0;
        }
        if (thr_out && g == 0) thr_out[0] = __int_as_float(0x7fc00000);
        cluster.sync();
        return;
    }
    // torch.quantile: rank = q * (cnt - 1) evaluated in fp32, lerp between floor and ceil ranks
    const float rank = q * static_cast<float>(cnt - 1ull);
    const float rank_lo_f = floorf(rank);
    const unsigned long long k_lo = static_cast<unsigned long long>(rank_lo_f);
    const unsigned long long k_hi = static_cast<unsigned long long>(ceilf(rank));
    const float w = rank - rank_lo_f;

    uint32_t prefix = 0u;
    unsigned int r = static_cast<unsigned int>(k_lo);
    unsigned int* leader_hist = cluster.map_shared_rank(&ghist[0][0], 0);
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        const uint32_t hi_mask = (pass == 0) ? 0u : (0xffffffffu << (shift + 8));
#pragma unroll
        for (int j = 0; j < kQPer; ++j) {
            if (j < per) {  // uniform: the warp collectives of hist_add see whole warps
                const bool active = ((valid >> j) & 1u) && ((key[j] & hi_mask) == prefix);
                hist_add(hist, (key[j] >> shift) & 0xffu, active);
            }
        }
        __syncthreads();
        if (tid < 256 && hist[tid]) atomicAdd(&leader_hist[pass * 256 + tid], hist[tid]);
        cluster.sync();
        if (tid < 256) {
            const unsigned int h = leader_hist[pass * 256 + tid];
            unsigned int inc = h;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
                if ((tid & 31) >= o) inc += up;
            }
            if ((tid & 31) == 31) s_wtot[tid >> 5] = inc;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            unsigned int base = 0;
            for (int wv = 0; wv < (tid >> 5); ++wv) base += s_wtot[wv];
            const unsigned int excl = base + inc - h;
            if (h != 0u && r >= excl && r < excl + h) {
                s_rank = r - excl;
                s_prefix = prefix | (static_cast<uint32_t>(tid) << shift);
            }
        }
        __syncthreads();
        r = s_rank;
        prefix = s_prefix;
        __syncthreads();
    }
    const uint32_t key_lo = prefix;
    // ---- the next order statistic: equal to key_lo if duplicated far enough, else the smallest key above it
    unsigned long long le = 0ull, mn = 0xffffffffull;
#pragma unroll
    for (int j = 0; j < kQPer; ++j) {
        if ((valid >> j) & 1u) {
            if (key[j] <= key_lo) ++le;
            else mn = min(mn, static_cast<unsigned long long>(key[j]));
        }
    }
    cluster_xchg(cluster, xq[1], le, mn, s_red);
    unsigned long long cnt_le = 0ull, min_gt = 0xffffffffull;
#pragma unroll
    for (int rr = 0; rr < kQCtas; ++rr) {
        cnt_le += xq[1][rr][0];
        min_gt = min(min_gt, xq[1][rr][1]);
    }
    const float x_lo = key_to_float(key_lo);
    const float x_hi = (k_hi == k_lo || cnt_le > k_hi) ? x_lo : key_to_float(static_cast<uint32_t>(min_gt));
    // at::lerp: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
    const float diff = x_hi - x_lo;
    float thr = (w < 0.5f) ? fmaf(w, diff, x_lo) : x_hi - diff * (1.f - w);
    if (no_nan == 0ull) thr = __int_as_float(0x7fc00000);
    if (thr_out && g == 0) thr_out[0] = thr;
    // (entropies * mask) >= threshold, and not padding (:361-364)
#pragma unroll
    for (int j = 0; j < kQPer; ++j) {
        const int i = j * kStride + g;
        if (j < per && i < n) out[i] = (((valid >> j) & 1u) && (key_to_float(key[j]) * 1.0f >= thr)) ? 1 : 0;
    }
    cluster.sync();  // the leader's histogram stays mapped until every CTA has read it
}

// rows_per_batch == 0: row r at r * row_stride; else row r = (b, t) at b * batch_stride + t * row_stride (a dlogits buffer
// that mirrors a strided logits view, ops.alloc_dlogits)
template <typename T>
__global__ void rescale_kernel(T* buf, int64_t n_rows, int64_t vocab, int64_t row_stride, int64_t rows_per_batch,
                               int64_t batch_stride, const float* actual, float expected) {
    const float a = actual[0];
    if (a == expected) return;  // the common case: nothing to do, the launch costs a few microseconds
    const float f = a / expected;
    const int64_t total = n_rows * vocab;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / vocab, c = i - r * vocab;
        int64_t off = r * row_stride;
        if (rows_per_batch > 0) {
            const int64_t b = r / rows_per_batch;
            off = b * batch_stride + (r - b * rows_per_batch) * row_stride;
        }
        T* p = buf + off + c;
        ElemTraits<T>::store(p, ElemTraits<T>::load(p) * f);
    }
}

}  // namespace
}  // namespace b200trl

using namespace b200trl;

extern "C" int64_t b200trl_entropy_quantile_workspace_bytes(int64_t) { return static_cast<int64_t>(sizeof(QuantileWs)); }

extern "C" int b200trl_entropy_quantile_mask(const float* entropies, const int32_t* mask, int64_t n, float threshold,
                                             void* workspace, uint8_t* out_mask, float* out_threshold,
                                             b200trl_stream_t stream) {
    B200TRL_REQUIRE(entropies && mask && out_mask && workspace, B200TRL_E_INVALID,
                    "entropy_quantile_mask: null pointer");
    B200TRL_REQUIRE(n > 0 && n < (int64_t(1) << 32), B200TRL_E_INVALID, "entropy_quantile_mask: bad size %lld",
                    (long long)n);
    B200TRL_REQUIRE(threshold >= 0.f && threshold <= 1.f, B200TRL_E_INVALID,
                    "entropy_quantile_mask: quantile %f outside [0,1]", threshold);
    static const bool use_cluster = !(getenv("B200TRL_QUANTILE_CLUSTER") && atoi(getenv("B200TRL_QUANTILE_CLUSTER")) == 0);
    if (use_cluster && n <= static_cast<int64_t>(kQCtas) * kSelBlock * kQPer) {
        entropy_quantile_cluster_kernel<<<kQCtas, kSelBlock, 0, as_stream(stream)>>>(
            entropies, mask, static_cast<int>(n), threshold, out_mask, out_threshold);
        return check_launch("entropy_quantile_cluster_kernel");
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entropy_quantile_kernel, kSelBlock, 0) != cudaSuccess ||
        per_sm < 1) {
        set_error("entropy_quantile_mask: occupancy query failed");
        return B200TRL_E_LAUNCH;
    }
    const int64_t want = (n + kSelBlock * 32 - 1) / (kSelBlock * 32);  // ~4 elements per thread per pass
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(per_sm) * num_sms()));
    QuantileWs* ws = static_cast<QuantileWs*>(workspace);
    void* params[] = {&entropies, &mask, &n, &threshold, &ws, &out_mask, &out_threshold};
    const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(entropy_quantile_kernel),
                                                     dim3(static_cast<unsigned>(grid)), dim3(kSelBlock), params, 0,
                                                     as_stream(stream));
    if (e != cudaSuccess) {
        set_error("entropy_quantile_mask: cooperative launch failed: %s", cudaGetErrorString(e));
        return B200TRL_E_LAUNCH;
    }
    return check_launch("entropy_quantile_kernel");
}

static int rescale_impl(void* buf, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride, int64_t rows_per_batch,
                        int64_t batch_stride, const float* actual, float expected, b200trl_stream_t stream);

extern "C" int b200trl_rescale_if_needed(void* buf, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                                         const float* actual, float expected, b200trl_stream_t stream) {
    return rescale_impl(buf, dtype, n_rows, vocab, row_stride, 0, 0, actual, expected, stream);
}

extern "C" int b200trl_rescale_if_needed_batched(void* buf, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride,
                                                 int64_t rows_per_batch, int64_t batch_stride, const float* actual,
                                                 float expected, b200trl_stream_t stream) {
    B200TRL_REQUIRE(rows_per_batch >= 0 && (rows_per_batch == 0 || n_rows % rows_per_batch == 0), B200TRL_E_INVALID,
                    "rescale_if_needed: rows_per_batch %lld does not divide n_rows %lld", (long long)rows_per_batch,
                    (long long)n_rows);
    return rescale_impl(buf, dtype, n_rows, vocab, row_stride, rows_per_batch, batch_stride, actual, expected, stream);
}

static int rescale_impl(void* buf, int dtype, int64_t n_rows, int64_t vocab, int64_t row_stride, int64_t rows_per_batch,
                        int64_t batch_stride, const float* actual, float expected, b200trl_stream_t stream) {
    B200TRL_REQUIRE(buf && actual, B200TRL_E_INVALID, "rescale_if_needed: null pointer");
    B200TRL_REQUIRE(expected != 0.f, B200TRL_E_INVALID, "rescale_if_needed: expected scale is zero");
    const int block = 256;
    const unsigned grid = static_cast<unsigned>(num_sms() * 8);
    cudaStream_t s = as_stream(stream);
    switch (dtype) {
        case B200TRL_BF16:
            rescale_kernel<<<grid, block, 0, s>>>(static_cast<__nv_bfloat16*>(buf), n_rows, vocab, row_stride, rows_per_batch,
                                                  batch_stride, actual, expected);
            break;
        case B200TRL_F16:
            rescale_kernel<<<grid, block, 0, s>>>(static_cast<__half*>(buf), n_rows, vocab, row_stride, rows_per_batch,
                                                  batch_stride, actual, expected);
            break;
        case B200TRL_F32:
            rescale_kernel<<<grid, block, 0, s>>>(static_cast<float*>(buf), n_rows, vocab, row_stride, rows_per_batch,
                                                  batch_stride, actual, expected);
            break;
        case B200TRL_F64:
            rescale_kernel<<<grid, block, 0, s>>>(static_cast<double*>(buf), n_rows, vocab, row_stride, rows_per_batch,
                                                  batch_stride, actual, expected);
            break;
        default: set_error("rescale_if_needed: unknown dtype %d", dtype); return B200TRL_E_UNSUPPORTED;
    }
    return check_launch("rescale_kernel");
}
