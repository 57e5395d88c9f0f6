// K1 "row" implementation: one CTA per logit row, plain (vectorised where aligned) global loads.
// Handles every dtype / vocab / stride.  For fwd+bwd the row is read twice by the same CTA; the second
// read is normally served by the 126 MB L2 (1 CTA of 1024 threads per SM keeps <= 148 rows in flight).
// The resident kernel (k1_resident.cu) is the single-HBM-pass product path for bf16.
//
// Replaces trl/trainer/utils.py:1430-1490 (+ autograd backward) — see include/b200trl.h.
#include "k1_args.cuh"

namespace b200trl {
namespace {

template <typename T>
struct Vec {
    static constexpr int N = 16 / sizeof(T);
};

// unpack a 16-byte vector of T into floats
template <typename T>
__device__ __forceinline__ void unpack16(const uint4& v, float* out);
template <>
__device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4& v, float* out) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        out[2 * i] = __uint_as_float(w[i] << 16);
        out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <>
__device__ __forceinline__ void unpack16<__half>(const uint4& v, float* out) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        out[2 * i] = f.x;
        out[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void unpack16<float>(const uint4& v, float* out) {
    out[0] = __uint_as_float(v.x);
    out[1] = __uint_as_float(v.y);
    out[2] = __uint_as_float(v.z);
    out[3] = __uint_as_float(v.w);
}
template <>
__device__ __forceinline__ void unpack16<double>(const uint4& v, float* out) {
    const double* d = reinterpret_cast<const double*>(&v);
    out[0] = static_cast<float>(d[0]);
    out[1] = static_cast<float>(d[1]);
}

template <typename T>
__device__ __forceinline__ uint4 pack16(const float* in);
template <>
__device__ __forceinline__ uint4 pack16<__nv_bfloat16>(const float* in) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <>
__device__ __forceinline__ uint4 pack16<__half>(const float* in) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 h = __floats2half2_rn(in[2 * i], in[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <>
__device__ __forceinline__ uint4 pack16<float>(const float* in) {
    return make_uint4(__float_as_uint(in[0]), __float_as_uint(in[1]), __float_as_uint(in[2]), __float_as_uint(in[3]));
}
template <>
__device__ __forceinline__ uint4 pack16<double>(const float* in) {
    uint4 v;
    double* d = reinterpret_cast<double*>(&v);
    d[0] = in[0];
    d[1] = in[1];
    return v;
}

// fold N raw logits into a partial with at most one rescale.  The exponent argument is a single FMA
// x*c - m (one rounding of a small number) — rounding x*c first would cost ~|y| * 2^-24 in the exponent,
// which peaked rows turn into a 1e-5-level entropy error.
template <int N>
__device__ __forceinline__ void partial_add(Partial& p, const float* x, float c) {
    float mx = x[0];
#pragma unroll
    for (int i = 1; i < N; ++i) mx = fmaxf(mx, x[i]);
    mx *= c;
    if (mx > p.m) {  // move the reference point (also the first touch: p.m == kNegBig)
        const float d = p.m - mx;
        const float f = ex2(d);
        p.u = f * fmaf(d, p.s, p.u);
        p.s *= f;
        p.m = mx;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float d = fmaf(x[i], c, -p.m);
        const float e = ex2(d);
        p.s += e;
        p.u = fmaf(e, d, p.u);
    }
}

template <int BLOCK>
__device__ __forceinline__ Partial block_reduce_partial(Partial p, Partial* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    p = partial_warp_reduce(p);
    if (lane == 0) smem[warp] = p;
    __syncthreads();
    Partial r = partial_empty();
    if (lane < BLOCK / 32) r = smem[lane];
    r = partial_warp_reduce(r);  // every warp computes the same tree => identical result in all threads
    return r;
}

// 1024-thread CTAs must stay at <= 32 registers so that two of them share an SM (a 52-register build measured 26 %
// slower forward-only at config 2: one CTA per SM cannot hide the HBM latency of its own row)
template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK == 1024 ? 2 : 1) k1_row_kernel(const K1Args a) {
    constexpr int VN = Vec<T>::N;
    __shared__ Partial s_part[32];
    __shared__ RowScalars s_row;
    __shared__ float s_ppo_count;

    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x;
    const T* x = reinterpret_cast<const T*>(a.logits) + logits_offset(a, row);
    const int64_t V = a.vocab;

    if (a.row_mask && !a.dlogits && a.row_mask[row] == 0) {  // masked forward: the row is not read, outputs are 0
        if (tid == 0) {
            if (a.logp) a.logp[row] = 0.f;
            if (a.entropy) a.entropy[row] = 0.f;
            if (a.lse) a.lse[row] = 0.f;
        }
        return;
    }
    if (a.skip_masked && a.dlogits && a.lse_in == nullptr && row_is_masked(a, row)) {
        // opt-in: a row the loss ignores is not read; outputs are zero (PPO: INVALID_LOGPROB), dlogits are zero
        T* dz = reinterpret_cast<T*>(a.dlogits) + dlogits_offset(a, row);
        for (int64_t j = tid; j < V; j += BLOCK) ElemTraits<T>::store(dz + j, 0.f);
        if (tid == 0) {
            if (a.logp) a.logp[row] = (a.gmode == G_PPO) ? 1.0f : 0.f;
            if (a.entropy) a.entropy[row] = 0.f;
            if (a.lse) a.lse[row] = 0.f;
        }
        return;
    }
    if (a.gmode == G_PPO && tid < 32) {
        const float n = ppo_unpadded_count(a, tid);
        if (tid == 0) s_ppo_count = n;
    }
    if (a.gmode == G_PPO) __syncthreads();
    if (tid == 0) s_row = load_row_scalars<T>(a, row, a.gmode == G_PPO ? s_ppo_count : 1.f);

    // split the row into [0,head) scalar, [head, head + nvec*VN) 16-byte vectors, tail scalar
    const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
    int64_t head = ((16 - (addr & 15)) & 15) / sizeof(T);
    if (head > V) head = V;
    const int64_t nvec = (V - head) / VN;
    const int64_t tail0 = head + nvec * VN;
    const uint4* xv = reinterpret_cast<const uint4*>(x + head);

    float lse2;
    float logp;
    if (a.lse_in == nullptr) {
        Partial p = partial_empty();
        for (int64_t i = tid; i < nvec; i += BLOCK) {
            float f[VN];
            unpack16<T>(__ldg(xv + i), f);
            partial_add<VN>(p, f, a.c);
        }
        for (int64_t i = tid; i < head + (V - tail0); i += BLOCK) {
            const int64_t j = (i < head) ? i : tail0 + (i - head);
            const float y = ElemTraits<T>::load(x + j);
            partial_add<1>(p, &y, a.c);
        }
        p = block_reduce_partial<BLOCK>(p, s_part);  // contains a __syncthreads: s_row is visible after it
        const RowStats rs = finish_row(p, s_row.x_sel, a.c);
        lse2 = rs.lse2;
        logp = rs.logp;
        if (tid == 0) {
            const bool pad = (a.gmode == G_PPO) && s_row.pad != 0.f;
            if (a.logp) a.logp[row] = pad ? 1.0f : rs.logp;  // INVALID_LOGPROB (ppo_trainer.py:81, 561-563)
            if (a.entropy) a.entropy[row] = rs.entropy;
            if (a.lse) a.lse[row] = rs.lse;
        }
    } else {
        __syncthreads();
        lse2 = a.lse_in[row] * kLog2e;
        logp = fmaf(s_row.x_sel, a.c, -lse2) * kLn2;
    }
    if (a.dlogits == nullptr) return;

    // ---- backward: dlogits = g' * (onehot - softmax), g' = g * inv_T
    const float gp = token_grad(a, s_row, logp) * a.inv_temp;
    T* dx = reinterpret_cast<T*>(a.dlogits) + dlogits_offset(a, row);
    const int64_t id = s_row.id;
    const bool same_align = ((reinterpret_cast<uintptr_t>(dx) & 15) == (addr & 15));
    if (same_align) {
        uint4* dv = reinterpret_cast<uint4*>(dx + head);
        for (int64_t i = tid; i < nvec; i += BLOCK) {
            float f[VN];
            if (gp != 0.f) {
                unpack16<T>(__ldg(xv + i), f);
                const int64_t e0 = head + i * VN;
#pragma unroll
                for (int k = 0; k < VN; ++k) {
                    const float pr = ex2(fmaf(f[k], a.c, -lse2));
                    f[k] = fmaf(-pr, gp, (e0 + k == id) ? gp : 0.f);
                }
            } else {
#pragma unroll
                for (int k = 0; k < VN; ++k) f[k] = 0.f;
            }
            dv[i] = pack16<T>(f);
        }
        for (int64_t i = tid; i < head + (V - tail0); i += BLOCK) {
            const int64_t j = (i < head) ? i : tail0 + (i - head);
            float o = 0.f;
            if (gp != 0.f) {
                const float pr = ex2(fmaf(ElemTraits<T>::load(x + j), a.c, -lse2));
                o = fmaf(-pr, gp, (j == id) ? gp : 0.f);
            }
            ElemTraits<T>::store(dx + j, o);
        }
    } else {
        for (int64_t j = tid; j < V; j += BLOCK) {
            float o = 0.f;
            if (gp != 0.f) {
                const float pr = ex2(fmaf(ElemTraits<T>::load(x + j), a.c, -lse2));
                o = fmaf(-pr, gp, (j == id) ? gp : 0.f);
            }
            ElemTraits<T>::store(dx + j, o);
        }
    }
}

template <typename T>
int launch_typed(const K1Args& a, cudaStream_t stream) {
    const dim3 grid(static_cast<unsigned>(a.n_rows));
    if (a.vocab >= 16384) {
        k1_row_kernel<T, 1024><<<grid, 1024, 0, stream>>>(a);
    } else {
        k1_row_kernel<T, 256><<<grid, 256, 0, stream>>>(a);
    }
    return check_launch("k1_row_kernel");
}

}  // namespace

int launch_k1_row(const K1Args& a, int dtype, cudaStream_t stream) {
    if (a.n_rows == 0) return B200TRL_OK;
    B200TRL_REQUIRE(a.n_rows <= 0x7fffffff, B200TRL_E_INVALID, "k1: n_rows %lld exceeds grid limit", (long long)a.n_rows);
    switch (dtype) {
        case B200TRL_BF16: return launch_typed<__nv_bfloat16>(a, stream);
        case B200TRL_F16: return launch_typed<__half>(a, stream);
        case B200TRL_F32: return launch_typed<float>(a, stream);
        case B200TRL_F64: return launch_typed<double>(a, stream);
        default: set_error("k1: unknown dtype %d", dtype); return B200TRL_E_UNSUPPORTED;
    }
}

}  // namespace b200trl
