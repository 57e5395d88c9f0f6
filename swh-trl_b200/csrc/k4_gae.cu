// K4: PPO reward shaping + reverse GAE + masked whitening in ONE cooperative launch.
//   ppo_trainer.py:500-506  pad fills        :510-516  KL reward + score at the last token
//   ppo_trainer.py:519-521  reward whitening :523-535  GAE reverse recurrence, returns, advantage whitening
//   trl/core.py:43-76       masked_mean / masked_var / masked_whiten (also exported stand-alone)
// The reference runs T sequential steps of ~5 tiny kernels (~2500 launches at B=64, T=512).  Here every row's
// recurrence A_t = delta_t + gamma*lam*A_{t+1} is a warp scan over affine maps: each lane folds a contiguous
// segment, a 5-step shuffle scan composes the 32 segment maps, then each lane replays its segment with the
// true carry-in.  Global masked statistics use per-CTA partials + grid.sync() and are folded in CTA order in
// double precision, so results are deterministic.
//
// Batches of up to 131 072 tokens (config 3 is 32 768) do not need the grid at all: ppo_gae_cluster_kernel keeps
// rewards / values / advantages of the whole batch in the shared memory of ONE 8-CTA thread-block cluster, reads every
// input once and writes every output once, and replaces the four to seven grid.sync() rounds (3-5 us each, 31-41 us
// in total at config 3) by cluster barriers with the partial sums exchanged through distributed shared memory.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200trl {
namespace {

constexpr int kBlock = 256;
constexpr int kMaxGrid = 1024;

struct GridWs {
    double part[kMaxGrid][2];
};

// sum of (a, b) over the whole grid; every thread returns the same totals.  slot selects a disjoint partial
// array so back-to-back reductions need one grid.sync each.
__device__ __forceinline__ void grid_sum2(cg::grid_group& grid, GridWs* ws, int slot, float a, float b, double& ta,
                                          double& tb, float* red) {
    float v[2] = {a, b};
    block_sum<2, kBlock>(v, red);
    GridWs* w = ws + slot;
    if (threadIdx.x == 0) {
        w->part[blockIdx.x][0] = static_cast<double>(v[0]);
        w->part[blockIdx.x][1] = static_cast<double>(v[1]);
    }
    __threadfence();
    grid.sync();
    // every warp folds the per-CTA partials the same way -- lane-strided sums, then a fixed shuffle tree -- so all
    // threads of the grid get bit-identical totals; a serial walk by every thread costs gridDim.x dependent L2 reads
    double sa = 0.0, sb = 0.0;
    const volatile double* p = &w->part[0][0];
    for (unsigned i = threadIdx.x & 31; i < gridDim.x; i += 32) {
        sa += p[2 * i];
        sb += p[2 * i + 1];
    }
    ta = warp_sum(sa);
    tb = warp_sum(sb);
}

// masked_whiten over the whole [n] array (trl/core.py:70-76); values are rewritten in place.
// mean = sum(x*m)/sum(m); var = sum((x-mean)^2*m)/sum(m) * n/(n-1); out = (x-mean)*rsqrt(var+1e-8) (+mean).
template <typename MaskFn>
__device__ __forceinline__ void grid_whiten(cg::grid_group& grid, GridWs* ws, int slot0, float* x, int64_t n, MaskFn keep,
                                            bool shift_mean, bool zero_masked, float* red, float* stats_out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kBlock;
    const int64_t start = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x;
    float s = 0.f, c = 0.f;
    for (int64_t i = start; i < n; i += stride) {
        const float k = keep(i) ? 1.f : 0.f;
        s += x[i] * k;
        c += k;
    }
    double ts, tc;
    grid_sum2(grid, ws, slot0, s, c, ts, tc, red);
    const float mean = static_cast<float>(ts / tc);
    float q = 0.f;
    for (int64_t i = start; i < n; i += stride) {
        const float d = x[i] - mean;
        q += keep(i) ? d * d : 0.f;
    }
    double tq, unused;
    grid_sum2(grid, ws, slot0 + 1, q, 0.f, tq, unused, red);
    const float var = static_cast<float>((tq / tc) * (tc / (tc - 1.0)));
    const float inv = rsqrtf(var + 1e-8f);
    for (int64_t i = start; i < n; i += stride) {
        float o = (x[i] - mean) * inv;
        if (!shift_mean) o += mean;
        if (zero_masked && !keep(i)) o = 0.f;
        x[i] = o;
    }
    if (stats_out && blockIdx.x == 0 && threadIdx.x == 0) {
        stats_out[0] = mean;
        stats_out[1] = var;
        stats_out[2] = static_cast<float>(tc);
    }
}

struct GaeArgs {
    const float* lp;
    const float* rlp;
    const float* values;
    const float* scores;
    const int64_t* seq_len;
    int64_t B, T;
    float kl_coef;
    int estimator;
    float gamma, lam;
    int whiten_rewards;
    GridWs* ws;
    float* rewards;
    float* adv;
    float* returns;
    float* lp_f;
    float* rlp_f;
    float* val_f;
};

__global__ void __launch_bounds__(kBlock) ppo_gae_kernel(const GaeArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float red[2 * 32];
    const int tid = threadIdx.x;
    const int64_t T = a.T, n = a.B * a.T;
    float* vals = a.val_f ? a.val_f : a.returns;  // filled values are parked in `returns` until the end

    // ---- phase 1: pad fills, KL reward, score at actual_end  (:500-516)
    for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
        const int64_t len = a.seq_len[b];
        const int64_t end = (len + 1 < T) ? len + 1 : len;  // :515
        const float score = a.scores[b];
        for (int64_t t = tid; t < T; t += kBlock) {
            const int64_t i = b * T + t;
            const bool pad = t > len, pad1 = t > len + 1;
            const float lp = pad ? 1.0f : a.lp[i];    // INVALID_LOGPROB (:502)
            const float rlp = pad ? 1.0f : a.rlp[i];  // :503
            const float v = pad1 ? 0.f : a.values[i];  // :506
            const float logr = rlp - lp;               // :510
            const float kl = (a.estimator == B200TRL_KL_K1) ? -logr : (expf(logr) - 1.f) - logr;  // :511
            float r = -a.kl_coef * kl;                 // :512
            if (t == end) r += score;                  // :516
            a.rewards[i] = r;
            vals[i] = v;
            if (a.lp_f) a.lp_f[i] = lp;
            if (a.rlp_f) a.rlp_f[i] = rlp;
        }
    }
    const int64_t* sl = a.seq_len;
    auto keep_p1 = [sl, T](int64_t i) { return (i % T) <= sl[i / T] + 1; };  // ~padding_mask_p1
    auto keep_p0 = [sl, T](int64_t i) { return (i % T) <= sl[i / T]; };      // ~padding_mask
    if (a.whiten_rewards) {  // :519-521
        __threadfence();
        grid.sync();
        grid_whiten(grid, a.ws, 0, a.rewards, n, keep_p1, /*shift_mean=*/false, /*zero_masked=*/true, red, nullptr);
    }
    __threadfence();
    grid.sync();

    // ---- phase 2: reverse GAE, one warp per row  (:523-533)
    {
        const int lane = tid & 31;
        const int64_t gwarp = static_cast<int64_t>(blockIdx.x) * (kBlock / 32) + (tid >> 5);
        const int64_t nwarp = static_cast<int64_t>(gridDim.x) * (kBlock / 32);
        const float k = a.gamma * a.lam;
        const int64_t L = (T + 31) / 32;  // segment length per lane
        for (int64_t b = gwarp; b < a.B; b += nwarp) {
            const float* r = a.rewards + b * T;
            const float* v = vals + b * T;
            const int64_t t0 = lane * L, t1 = min(T, t0 + L);
            // local fold with zero carry-in: A(t0) = Q + P * carry
            float Q = 0.f, P = 1.f;
            for (int64_t t = t1 - 1; t >= t0; --t) {
                const float nv = (t + 1 < T) ? v[t + 1] : 0.f;
                const float delta = r[t] + a.gamma * nv - v[t];
                Q = delta + k * Q;
                P *= k;
            }
            // suffix scan over lanes: carry_l = A at the first step of lane l+1
            float cQ = Q, cP = P;  // composite map of lanes [l, l+o)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float nQ = __shfl_down_sync(0xffffffffu, cQ, o);
                const float nP = __shfl_down_sync(0xffffffffu, cP, o);
                if (lane + o < 32) {
                    cQ = cQ + cP * nQ;
                    cP = cP * nP;
                }
            }
            float carry = __shfl_down_sync(0xffffffffu, cQ, 1);  // A at start of the next lane's segment
            if (lane == 31) carry = 0.f;
            float A = carry;
            for (int64_t t = t1 - 1; t >= t0; --t) {
                const float nv = (t + 1 < T) ? v[t + 1] : 0.f;
                const float delta = r[t] + a.gamma * nv - v[t];
                A = delta + k * A;
                a.adv[b * T + t] = A;
            }
        }
    }
    __threadfence();
    grid.sync();
    // returns = advantages + values (:533); must read vals before it is overwritten when parked in `returns`
    {
        const int64_t stride = static_cast<int64_t>(gridDim.x) * kBlock;
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * kBlock + tid; i < n; i += stride)
            a.returns[i] = a.adv[i] + vals[i];
    }
    // ---- phase 3: advantage whitening, pads to zero (:534-535)
    grid_whiten(grid, a.ws, 2, a.adv, n, keep_p0, /*shift_mean=*/true, /*zero_masked=*/true, red, nullptr);
}


// ------------------------------------------------------------------ single-cluster variant (small batches)
constexpr int kClusterCtas = 8;
constexpr int kClBlock = 1024;
constexpr int kClMaxElems = 16384;  // per CTA: 3 fp32 arrays of this many elements = 192 KB of shared memory

__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f64(const double* local, uint32_t rank, double v) {
    uint32_t addr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                 : "=r"(addr)
                 : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(local))), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// sum of (a, b) over the cluster: block tree, every CTA posts its partial into every CTA's exchange row (DSMEM),
// one cluster barrier, then everybody adds the 8 partials in rank order -> bit-identical totals in all threads
__device__ __forceinline__ void cluster_sum2(double (*xch)[2], float a, float b, double& ta, double& tb, float* red) {
    float v[2] = {a, b};
    block_sum<2, kClBlock>(v, red);
    const uint32_t me = cl_rank();
    if (threadIdx.x < kClusterCtas) {
        st_cluster_f64(&xch[me][0], threadIdx.x, static_cast<double>(v[0]));
        st_cluster_f64(&xch[me][1], threadIdx.x, static_cast<double>(v[1]));
    }
    cl_sync();
    ta = 0.0;
    tb = 0.0;
#pragma unroll
    for (int r = 0; r < kClusterCtas; ++r) {
        ta += xch[r][0];
        tb += xch[r][1];
    }
}

// masked_whiten of x[0, n_loc) (shared memory) with statistics over the whole cluster.  Every pass of the kernel
// walks the CTA's elements with the same stride, so element tid + j * kClBlock always belongs to thread tid: its mask
// is bit j of `bits` (computed once in phase 1, where the sequence length is loaded anyway -- the per-element i / T,
// i % T and seq_len loads of a mask functor were a measurable part of a 13 us kernel).
__device__ __forceinline__ void cluster_whiten(double (*xch)[kClusterCtas][2], int slot0, float* x, int n_loc, uint32_t bits,
                                               bool shift_mean, float* red) {
    const int tid = threadIdx.x;
    auto keep = [bits, tid](int i) { return ((bits >> ((i - tid) / kClBlock)) & 1u) != 0u; };
    float s = 0.f, c = 0.f;
    for (int i = tid; i < n_loc; i += kClBlock) {
        const float k = keep(i) ? 1.f : 0.f;
        s += x[i] * k;
        c += k;
    }
    double ts, tc, tq, unused;
    cluster_sum2(xch[slot0], s, c, ts, tc, red);
    const float mean = static_cast<float>(ts / tc);
    float q = 0.f;
    for (int i = tid; i < n_loc; i += kClBlock) {
        const float d = x[i] - mean;
        q += keep(i) ? d * d : 0.f;
    }
    cluster_sum2(xch[slot0 + 1], q, 0.f, tq, unused, red);
    const float var = static_cast<float>((tq / tc) * (tc / (tc - 1.0)));
    const float inv = rsqrtf(var + 1e-8f);
    for (int i = tid; i < n_loc; i += kClBlock) {
        float o = (x[i] - mean) * inv;
        if (!shift_mean) o += mean;
        if (!keep(i)) o = 0.f;
        x[i] = o;
    }
    __syncthreads();
}

__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClBlock, 1)
    ppo_gae_cluster_kernel(const GaeArgs a, const int rows_per_cta) {
    extern __shared__ __align__(16) float sm_dyn[];
    __shared__ float red[2 * 32];
    __shared__ double xch[4][kClusterCtas][2];  // one exchange row set per reduction: never reused within a launch
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = static_cast<int>(a.T);
    const int64_t row0 = static_cast<int64_t>(cl_rank()) * rows_per_cta;
    const int64_t row1 = (row0 + rows_per_cta < a.B) ? row0 + rows_per_cta : a.B;
    const int rows = row1 > row0 ? static_cast<int>(row1 - row0) : 0;
    const int n_loc = rows * T;
    float* r_s = sm_dyn;                                          // rewards
    float* v_s = sm_dyn + static_cast<size_t>(rows_per_cta) * T;  // filled values
    float* a_s = v_s + static_cast<size_t>(rows_per_cta) * T;     // advantages
    const int64_t g0 = row0 * T;                                  // global index of local element 0

    // ---- phase 1: pad fills, KL reward, score at actual_end  (:500-516)
    uint32_t keep_p0 = 0u, keep_p1 = 0u;  // bit j: element tid + j * kClBlock is not padding / not padding shifted by one
    for (int i = tid, j = 0; i < n_loc; i += kClBlock, ++j) {
        const int rl = i / T, t = i - rl * T;
        const int64_t b = row0 + rl, gi = g0 + i;
        const int64_t len = a.seq_len[b];
        const int64_t end = (len + 1 < T) ? len + 1 : len;  // :515
        const bool pad = t > len, pad1 = t > len + 1;
        keep_p0 |= pad ? 0u : (1u << j);
        keep_p1 |= pad1 ? 0u : (1u << j);
        // the three tensors are full [B, T]: load unconditionally (the loads then fly together with the seq_len load
        // instead of behind it) and select afterwards
        const float lp_raw = a.lp[gi], rlp_raw = a.rlp[gi], v_raw = a.values[gi];
        const float lp = pad ? 1.0f : lp_raw;        // INVALID_LOGPROB (:502)
        const float rlp = pad ? 1.0f : rlp_raw;      // :503
        const float v = pad1 ? 0.f : v_raw;          // :506
        const float logr = rlp - lp;                 // :510
        const float kl = (a.estimator == B200TRL_KL_K1) ? -logr : (expf(logr) - 1.f) - logr;  // :511
        float r = -a.kl_coef * kl;                   // :512
        if (t == end) r += a.scores[b];              // :516
        r_s[i] = r;
        v_s[i] = v;
        if (a.lp_f) a.lp_f[gi] = lp;
        if (a.rlp_f) a.rlp_f[gi] = rlp;
        if (a.val_f) a.val_f[gi] = v;
    }
    __syncthreads();
    if (a.whiten_rewards)  // :519-521   (keep_p1 = ~padding_mask_p1, keep_p0 = ~padding_mask)
        cluster_whiten(xch, 0, r_s, n_loc, keep_p1, /*shift_mean=*/false, red);
    for (int i = tid; i < n_loc; i += kClBlock) a.rewards[g0 + i] = r_s[i];

    // ---- phase 2: reverse GAE, one warp per row, 32 steps per round  (:523-533)
    // With few rows per CTA (config 3: 8 rows, 32 warps) a row is cut into W segments of whole 32-step chunks, one warp
    // each: every warp scans its segment with a zero carry-in, then the carries are chained over the W segments (A at a
    // segment's first step = its zero-carry value + k^len * carry-in) and added back as k^(steps to the segment's end)
    // * carry-in.  The serial part of a row drops from T / 32 chunk scans to T / (32 W) + W.
    const int n_chunks = (T + 31) / 32;
    int W = 1;
    while (W * 2 * rows <= kClBlock / 32 && W * 2 <= n_chunks && W < 8) W *= 2;
    if (W > 1) {
        const float k = a.gamma * a.lam;
        const float log2k = log2f(k);  // k in (0, 1]; k == 0 -> -inf -> powers are 0, as they should be
        const int seg_chunks = (n_chunks + W - 1) / W;
        float* seg_first = red;  // [rows * W] zero-carry advantage at each segment's first step (red: 64 floats, >= 32)
        const int rl = warp / W, sg = warp - rl * W;
        const bool active = rl < rows;
        const int c_lo = sg * seg_chunks, c_hi = min(c_lo + seg_chunks, n_chunks);  // chunks [c_lo, c_hi) of the row
        if (active) {
            const float* r = r_s + rl * T;
            const float* v = v_s + rl * T;
            float carry = 0.f;
            for (int c = c_hi - 1; c >= c_lo; --c) {
                const int t = c * 32 + lane;
                float Q = 0.f, P = 1.f;
                if (t < T) {
                    const float nv = (t + 1 < T) ? v[t + 1] : 0.f;
                    Q = r[t] + a.gamma * nv - v[t];
                    P = k;
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float nQ = __shfl_down_sync(0xffffffffu, Q, o);
                    const float nP = __shfl_down_sync(0xffffffffu, P, o);
                    if (lane + o < 32) {
                        Q = Q + P * nQ;
                        P = P * nP;
                    }
                }
                const float A = Q + P * carry;
                if (t < T) a_s[rl * T + t] = A;
                carry = __shfl_sync(0xffffffffu, A, 0);
            }
            if (lane == 0) seg_first[rl * W + sg] = (c_lo < c_hi) ? carry : 0.f;
        }
        __syncthreads();
        if (active && sg < W - 1 && c_lo < c_hi) {
            // carry into this segment = true advantage at the first step of the next segment
            float cin = 0.f;
            for (int q = W - 1; q > sg; --q) {
                const int q_lo = q * seg_chunks, q_hi = min(q_lo + seg_chunks, n_chunks);
                if (q_lo >= q_hi) continue;
                const int steps = min(q_hi * 32, T) - q_lo * 32;  // real steps of segment q
                cin = seg_first[rl * W + q] + exp2f(log2k * static_cast<float>(steps)) * cin;
            }
            const int end = min(c_hi * 32, T);  // one past the segment's last real step
            for (int t = c_lo * 32 + lane; t < end; t += 32)
                a_s[rl * T + t] += exp2f(log2k * static_cast<float>(end - t)) * cin;
        }
    } else {
        const float k = a.gamma * a.lam;
        for (int rl = warp; rl < rows; rl += kClBlock / 32) {
            const float* r = r_s + rl * T;
            const float* v = v_s + rl * T;
            float carry = 0.f;  // A_{t+1} of the chunk's last step
            for (int c0 = ((T - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
                const int t = c0 + lane;
                float Q = 0.f, P = 1.f;  // identity map for steps past the row's end
                if (t < T) {
                    const float nv = (t + 1 < T) ? v[t + 1] : 0.f;
                    Q = r[t] + a.gamma * nv - v[t];
                    P = k;
                }
                // reverse inclusive scan of the affine maps A -> Q + P * A over the 32 lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float nQ = __shfl_down_sync(0xffffffffu, Q, o);
                    const float nP = __shfl_down_sync(0xffffffffu, P, o);
                    if (lane + o < 32) {
                        Q = Q + P * nQ;
                        P = P * nP;
                    }
                }
                const float A = Q + P * carry;
                if (t < T) a_s[rl * T + t] = A;
                carry = __shfl_sync(0xffffffffu, A, 0);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < n_loc; i += kClBlock) a.returns[g0 + i] = a_s[i] + v_s[i];  // :533
    // ---- phase 3: advantage whitening, pads to zero (:534-535)
    cluster_whiten(xch, 2, a_s, n_loc, keep_p0, /*shift_mean=*/true, red);
    for (int i = tid; i < n_loc; i += kClBlock) a.adv[g0 + i] = a_s[i];
    cl_sync();  // no CTA leaves while a peer may still write into its exchange rows
}

int launch_gae_cluster(const GaeArgs& a, int rows_per_cta, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(3) * rows_per_cta * a.T * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(ppo_gae_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(3 * kClMaxElems * sizeof(float))) != cudaSuccess) {
            set_error("ppo_gae_cluster_kernel: cannot reserve shared memory");
            return B200TRL_E_LAUNCH;
        }
        configured = true;
    }
    ppo_gae_cluster_kernel<<<kClusterCtas, kClBlock, smem, stream>>>(a, rows_per_cta);
    return check_launch("ppo_gae_cluster_kernel");
}

struct WhitenArgs {
    const float* values;
    const uint8_t* mask;
    int64_t n;
    int shift_mean;
    GridWs* ws;
    float* out;
    float* stats;
};

__global__ void __launch_bounds__(kBlock) masked_whiten_kernel(const WhitenArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float red[2 * 32];
    const uint8_t* m = a.mask;
    auto keep = [m](int64_t i) { return m[i] != 0; };
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kBlock;
    const int64_t start = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x;
    if (a.out) {
        for (int64_t i = start; i < a.n; i += stride) a.out[i] = a.values[i];
        // each thread re-reads only what it wrote itself: no sync needed
        grid_whiten(grid, a.ws, 0, a.out, a.n, keep, a.shift_mean != 0, /*zero_masked=*/false, red, a.stats);
    } else {
        // statistics only (masked_mean / masked_var)
        float s = 0.f, c = 0.f;
        for (int64_t i = start; i < a.n; i += stride) {
            const float k = keep(i) ? 1.f : 0.f;
            s += a.values[i] * k;
            c += k;
        }
        double ts, tc, tq, unused;
        grid_sum2(grid, a.ws, 0, s, c, ts, tc, red);
        const float mean = static_cast<float>(ts / tc);
        float q = 0.f;
        for (int64_t i = start; i < a.n; i += stride) {
            const float d = a.values[i] - mean;
            q += keep(i) ? d * d : 0.f;
        }
        grid_sum2(grid, a.ws, 1, q, 0.f, tq, unused, red);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.stats[0] = mean;
            a.stats[1] = static_cast<float>((tq / tc) * (tc / (tc - 1.0)));
            a.stats[2] = static_cast<float>(tc);
        }
    }
}

template <typename Args, typename Kern>
int coop_launch(Kern kern, const Args& args, int64_t want_blocks, cudaStream_t stream, const char* name) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlock, 0) != cudaSuccess || per_sm < 1) {
        set_error("%s: occupancy query failed", name);
        return B200TRL_E_LAUNCH;
    }
    int64_t grid = std::min<int64_t>(want_blocks, std::min<int64_t>(static_cast<int64_t>(per_sm) * num_sms(), kMaxGrid));
    if (grid < 1) grid = 1;
    void* params[] = {const_cast<Args*>(&args)};
    const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(static_cast<unsigned>(grid)),
                                                     dim3(kBlock), params, 0, stream);
    if (e != cudaSuccess) {
        set_error("%s: cooperative launch failed: %s", name, cudaGetErrorString(e));
        return B200TRL_E_LAUNCH;
    }
    return check_launch(name);
}

}  // namespace
}  // namespace b200trl

using namespace b200trl;

extern "C" int64_t b200trl_ppo_gae_workspace_bytes(int64_t, int64_t) { return 4 * static_cast<int64_t>(sizeof(GridWs)); }
extern "C" int64_t b200trl_masked_workspace_bytes(int64_t) { return 2 * static_cast<int64_t>(sizeof(GridWs)); }

extern "C" int b200trl_ppo_rewards_gae(const float* logprobs, const float* ref_logprobs, const float* values,
                                       const float* scores, const int64_t* sequence_lengths, int64_t B, int64_t T,
                                       float kl_coef, int kl_estimator, float gamma, float lam, int whiten_rewards,
                                       void* workspace, float* rewards, float* advantages, float* returns,
                                       float* logprobs_f, float* ref_logprobs_f, float* values_f,
                                       b200trl_stream_t stream) {
    B200TRL_REQUIRE(logprobs && ref_logprobs && values && scores && sequence_lengths && workspace && rewards &&
                        advantages && returns,
                    B200TRL_E_INVALID, "ppo_rewards_gae: null pointer");
    B200TRL_REQUIRE(B > 0 && T > 0, B200TRL_E_INVALID, "ppo_rewards_gae: bad shape");
    B200TRL_REQUIRE(kl_estimator == B200TRL_KL_K1 || kl_estimator == B200TRL_KL_K3, B200TRL_E_INVALID,
                    "ppo_rewards_gae: unknown kl estimator %d", kl_estimator);
    GaeArgs a{logprobs, ref_logprobs, values, scores, sequence_lengths, B, T, kl_coef, kl_estimator, gamma, lam,
              whiten_rewards, static_cast<GridWs*>(workspace), rewards, advantages, returns, logprobs_f,
              ref_logprobs_f, values_f};
    // one 8-CTA cluster when the batch fits its shared memory (B200TRL_GAE_CLUSTER=0: always the cooperative grid)
    static const bool use_cluster = !(getenv("B200TRL_GAE_CLUSTER") && atoi(getenv("B200TRL_GAE_CLUSTER")) == 0);
    const int64_t rows_per_cta = (B + kClusterCtas - 1) / kClusterCtas;
    if (use_cluster && rows_per_cta * T <= kClMaxElems)
        return launch_gae_cluster(a, static_cast<int>(rows_per_cta), as_stream(stream));
    return coop_launch(ppo_gae_kernel, a, B, as_stream(stream), "ppo_gae_kernel");
}

extern "C" int b200trl_masked_whiten(const float* values, const uint8_t* mask, int64_t n, int shift_mean,
                                     void* workspace, float* out, float* stats, b200trl_stream_t stream) {
    B200TRL_REQUIRE(values && mask && workspace && (out || stats), B200TRL_E_INVALID, "masked_whiten: null pointer");
    B200TRL_REQUIRE(n > 0, B200TRL_E_INVALID, "masked_whiten: empty input");
    WhitenArgs a{values, mask, n, shift_mean, static_cast<GridWs*>(workspace), out, stats};
    return coop_launch(masked_whiten_kernel, a, (n + kBlock * 4 - 1) / (kBlock * 4), as_stream(stream),
                       "masked_whiten_kernel");
}
