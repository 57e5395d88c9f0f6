// Shared device/host helpers for libb200trl (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>

#include "../../include/b200trl.h"

namespace b200trl {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegBig = -1.0e30f;  // finite stand-in for -inf so that (m_old - m_new) * 0 stays 0

// ---------------------------------------------------------------- host side
void set_error(const char* fmt, ...);
const char* last_error();
int check_launch(const char* what);
int num_sms();

#define B200TRL_REQUIRE(cond, code, ...)   \
    do {                                   \
        if (!(cond)) {                     \
            ::b200trl::set_error(__VA_ARGS__); \
            return (code);                 \
        }                                  \
    } while (0)

inline cudaStream_t as_stream(b200trl_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------- device math
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Online-softmax partial in log2 units: reference point m, S = sum 2^(y-m), U = sum 2^(y-m) (y-m).
// lse2 = m + log2 S; entropy = ln2 * (log2 S - U / S).  Accumulating (y - m) instead of y avoids the
// cancellation of two O(|max logit|) terms when the distribution is peaked.
struct Partial {
    float m, s, u;
};

__device__ __forceinline__ Partial partial_empty() { return Partial{kNegBig, 0.f, 0.f}; }

__device__ __forceinline__ Partial partial_merge(const Partial& a, const Partial& b) {
    const float m = fmaxf(a.m, b.m);
    const float da = a.m - m, db = b.m - m;  // <= 0, finite (kNegBig is finite)
    const float fa = ex2(da), fb = ex2(db);
    Partial r;
    r.m = m;
    r.s = a.s * fa + b.s * fb;
    r.u = fa * fmaf(da, a.s, a.u) + fb * fmaf(db, b.s, b.u);
    return r;
}

__device__ __forceinline__ Partial partial_warp_reduce(Partial p) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Partial q;
        q.m = __shfl_xor_sync(0xffffffffu, p.m, o);
        q.s = __shfl_xor_sync(0xffffffffu, p.s, o);
        q.u = __shfl_xor_sync(0xffffffffu, p.u, o);
        p = partial_merge(p, q);
    }
    return p;
}

// Same result as the merge tree up to fp32 round-off, but one ex2 per lane instead of two per merge level: find the
// warp-wide reference first (5 max shuffles), rescale once, then plain shuffle sums.  This sits on the per-row
// critical path of the resident kernel (consumer warps -> reducer warp -> RowResult).
__device__ __forceinline__ Partial partial_warp_reduce_fast(Partial p) {
    float m = p.m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float d = p.m - m;  // <= 0, finite
    const float f = ex2(d);
    float s = p.s * f;
    float u = f * fmaf(d, p.s, p.u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        u += __shfl_xor_sync(0xffffffffu, u, o);
    }
    return Partial{m, s, u};
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum K per-thread values over a CTA of BLOCK threads (fixed tree => run-to-run deterministic); every thread
// gets the totals.  smem: K * 32 floats.
template <int K, int BLOCK>
__device__ __forceinline__ void block_sum(float (&v)[K], float* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) smem[k * 32 + warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(lane < BLOCK / 32 ? smem[k * 32 + lane] : 0.f);
    __syncthreads();
}

// Row results from a finished partial.  x_sel is the raw selected logit, c = inv_T * log2e.
struct RowStats {
    float lse;      // natural-log logsumexp of the temperature-scaled row
    float lse2;     // same in log2 units
    float entropy;  // nats
    float logp;     // log_softmax at the selected id
};
__device__ __forceinline__ RowStats finish_row(const Partial& p, float x_sel, float c) {
    RowStats r;
    const float l2s = log2f(p.s);  // full-precision log2: once per row
    r.lse2 = p.m + l2s;
    r.lse = r.lse2 * kLn2;
    r.entropy = kLn2 * (l2s - p.u / p.s);
    r.logp = fmaf(x_sel, c, -r.lse2) * kLn2;
    return r;
}

// ---------------------------------------------------------------- element types
template <typename T>
struct ElemTraits;
template <>
struct ElemTraits<__nv_bfloat16> {
    static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <>
struct ElemTraits<__half> {
    static __device__ __forceinline__ float load(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ void store(__half* p, float v) { *p = __float2half_rn(v); }
};
template <>
struct ElemTraits<float> {
    static __device__ __forceinline__ float load(const float* p) { return *p; }
    static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <>
struct ElemTraits<double> {
    static __device__ __forceinline__ float load(const double* p) { return static_cast<float>(*p); }
    static __device__ __forceinline__ void store(double* p, float v) { *p = static_cast<double>(v); }
};

inline size_t dtype_size(int dtype) {
    switch (dtype) {
        case B200TRL_BF16:
        case B200TRL_F16: return 2;
        case B200TRL_F32: return 4;
        case B200TRL_F64: return 8;
        default: return 0;
    }
}

}  // namespace b200trl
