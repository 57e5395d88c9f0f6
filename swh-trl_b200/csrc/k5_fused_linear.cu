// K5 (forward): lm_head GEMM fused with the online log-softmax statistics on tcgen05 / TMEM — the [N, V] logits
// never exist in memory.  SURVEY §8f-1, BASELINE config 4 (hidden 3584 -> V = 152064).
//
//   logits[r, v] = sum_k hidden[r, k] * W[v, k]           (bf16 x bf16 -> fp32 in tensor memory)
//   per row: logp[ids[r]], entropy, lse of logits * inv_T  (same (m, S, U) statistics as K1)
//
// One CTA per SM, persistent over work items (m-tile, vocabulary group):
//   warp 0   TMA producer: 2-D tiled bulk-tensor loads (SWIZZLE_128B) of a 128 x 64 hidden tile and a 256 x 64 weight
//            tile per k-block into a 4-stage shared-memory ring (mbarrier complete_tx)
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16) four times
//            per k-block; tcgen05.commit releases the stage / publishes the accumulator
//   warp 2   allocates / frees the 512 TMEM columns (two 128 x 256 fp32 accumulators, ping-pong)
//   warps 4-7  epilogue: tcgen05.ld 32 lanes x 32 columns at a time; thread == row; folds the tile into the row's
//            running (m, S, U), picks up the selected logit; the next tile's MMAs run meanwhile
// Row partials per vocabulary group go to a small workspace and a second tiny kernel merges them.
//
// This first version uses single-CTA MMAs without operand multicast, so it is L2->SMEM bound below the library
// GEMM; it exists to remove the logits round trip for the no-grad log-prob passes and as the base
// for the cta_group::2 + multicast version.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace b200trl {
namespace {

constexpr int kBM = 128, kBN = 256, kBK = 64, kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK * 2;  // 16 KB
constexpr int kBBytes = kBN * kBK * 2;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr float kSlack5 = 6.0f;

struct Bars {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t s2u(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(b)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tK5_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra K5_DONE;\n\tbra K5_WAIT;\n\tK5_DONE:\n\t}" ::"r"(s2u(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            s2u(dst)),
        "l"(map), "r"(x), "r"(y), "r"(s2u(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i of the warp holds lane base + i)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major operand tile written by TMA with SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64))
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;            // leading byte offset (unused for swizzled K-major) = 16 B
    d |= static_cast<uint64_t>(1024 >> 4) << 32;    // stride byte offset: 8 rows x 128 B
    d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D fp32, A/B bf16, both K-major, M=128, N=256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kBN >> 3) << 17) |
                            (static_cast<uint32_t>(kBM >> 4) << 24);

struct K5Args {
    int64_t n_rows, vocab, hidden;
    const int64_t* ids;
    float c;  // inv_T * log2(e)
    int n_mtiles, n_ntiles, n_groups, tiles_per_group;
    float4* partial;  // [n_groups][n_mtiles * 128] : (m, S, U, x_sel or NaN)
};

__global__ void __launch_bounds__(kThreads, 1)
k5_fwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const K5Args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment in the shared window: align by hand (1 KB of slack is allocated)
    unsigned char* smem = smem_raw + ((1024u - (s2u(smem_raw) & 1023u)) & 1023u);
    Bars& bars = *reinterpret_cast<Bars*>(smem + kStages * kStageBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.n_mtiles * a.n_groups;
    const int kblocks = static_cast<int>((a.hidden + kBK - 1) / kBK);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            bar_init(&bars.full[s], 1);
            bar_init(&bars.empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            bar_init(&bars.tmem_full[b], 1);
            bar_init(&bars.tmem_empty[b], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(&bars.tmem_base)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars.tmem_base);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int g = item / a.n_mtiles, mt = item - g * a.n_mtiles;
                const int nt0 = g * a.tiles_per_group, nt1 = min(a.n_ntiles, nt0 + a.tiles_per_group);
                for (int nt = nt0; nt < nt1; ++nt) {
                    for (int kb = 0; kb < kblocks; ++kb) {
                        bar_wait(&bars.empty[stage], phase ^ 1u);
                        unsigned char* sa = smem + stage * kStageBytes;
                        bar_expect_tx(&bars.full[stage], kStageBytes);
                        tma_load_2d(sa, &map_a, kb * kBK, mt * kBM, &bars.full[stage]);
                        tma_load_2d(sa + kABytes, &map_b, kb * kBK, nt * kBN, &bars.full[stage]);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int buf = 0;
            uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int g = item / a.n_mtiles;
                const int nt0 = g * a.tiles_per_group, nt1 = min(a.n_ntiles, nt0 + a.tiles_per_group);
                for (int nt = nt0; nt < nt1; ++nt) {
                    bar_wait(&bars.tmem_empty[buf], acc_phase ^ 1u);  // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * kBN);
                    for (int kb = 0; kb < kblocks; ++kb) {
                        bar_wait(&bars.full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = s2u(smem + stage * kStageBytes);
                        const uint64_t adesc = make_kmajor_sw128_desc(sa);
                        const uint64_t bdesc = make_kmajor_sw128_desc(sa + kABytes);
#pragma unroll
                        for (int k = 0; k < kBK / kUmmaK; ++k) {
                            // advance 32 B (16 bf16) inside the 128 B swizzle atom: +2 in the (>>4) address field
                            tc_mma(tmem_d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), kIdesc,
                                   (kb | k) != 0 ? 1u : 0u);
                        }
                        tc_commit(&bars.empty[stage]);  // the stage can be refilled once these MMAs have read it
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    tc_commit(&bars.tmem_full[buf]);  // accumulator complete -> epilogue
                    buf ^= 1;
                    if (buf == 0) acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue: thread == accumulator row
        const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
        const int row_in_tile = ew * 32 + lane;
        const float c = a.c;
        int buf = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int g = item / a.n_mtiles, mt = item - g * a.n_mtiles;
            const int nt0 = g * a.tiles_per_group, nt1 = min(a.n_ntiles, nt0 + a.tiles_per_group);
            const int64_t row = static_cast<int64_t>(mt) * kBM + row_in_tile;
            const int64_t id = (row < a.n_rows) ? a.ids[row] : -1;
            float m = kNegBig, S = 0.f, U = 0.f, xsel = __int_as_float(0x7fc00000);
            for (int nt = nt0; nt < nt1; ++nt) {
                bar_wait(&bars.tmem_full[buf], acc_phase);
                tc_fence_after();
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(buf * kBN);
                const int64_t col0 = static_cast<int64_t>(nt) * kBN;
#pragma unroll 1
                for (int j = 0; j < kBN / 32; ++j) {
                    float v[32];
                    tc_ld32(tbase + static_cast<uint32_t>(j * 32), v);
                    const int64_t cj = col0 + j * 32;
                    const int64_t left = a.vocab - cj;
                    const int valid = left < 32 ? static_cast<int>(left) : 32;  // columns past V are padding
                    if (valid <= 0) continue;
                    if (valid < 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i >= valid) v[i] = -INFINITY;  // 2^(-inf) = 0 and -inf never wins the max
                    }
                    float mx = v[0];
#pragma unroll
                    for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
                    mx *= c;
                    if (mx > m + kSlack5) {
                        const float d = m - mx, f = ex2(d);
                        U = f * fmaf(d, S, U);
                        S *= f;
                        m = mx;
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float d = fmaf(v[i], c, -m);
                        const float e = ex2(d);
                        S += e;
                        U = (i < valid || valid == 32) ? fmaf(e, d, U) : U;  // 0 * -inf would poison the entropy term
                    }
                    if (id >= cj && id < cj + 32) {
                        const int want = static_cast<int>(id - cj);
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i == want) xsel = v[i];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(&bars.tmem_empty[buf]);
                buf ^= 1;
                if (buf == 0) acc_phase ^= 1u;
            }
            if (row < a.n_rows)
                a.partial[static_cast<int64_t>(g) * a.n_mtiles * kBM + row] = make_float4(m, S, U, xsel);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

__global__ void k5_merge_kernel(const float4* __restrict__ partial, int n_groups, int64_t padded_rows, int64_t n_rows,
                                float c, float* __restrict__ logp, float* __restrict__ entropy,
                                float* __restrict__ lse) {
    const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (row >= n_rows) return;
    Partial tot = partial_empty();
    float xsel = __int_as_float(0x7fc00000);
    for (int g = 0; g < n_groups; ++g) {
        const float4 p = partial[static_cast<int64_t>(g) * padded_rows + row];
        tot = partial_merge(tot, Partial{p.x, p.y, p.z});
        if (!isnan(p.w)) xsel = p.w;
    }
    const RowStats st = finish_row(tot, xsel, c);
    if (logp) logp[row] = st.logp;
    if (entropy) entropy[row] = st.entropy;
    if (lse) lse[row] = st.lse;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t row_stride_elems, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("fused_linear: cuTensorMapEncodeTiled is not available from the driver");
        return B200TRL_E_LAUNCH;
    }
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(row_stride_elems) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("fused_linear: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
        return B200TRL_E_INVALID;
    }
    return B200TRL_OK;
}

int env_int5(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

int groups_for(int64_t n_rows, int64_t vocab) {
    const int n_mtiles = static_cast<int>((n_rows + kBM - 1) / kBM);
    const int n_ntiles = static_cast<int>((vocab + kBN - 1) / kBN);
    // static round-robin over items: >= 16 waves keeps the ragged last wave under ~6 % (4 waves cost 14 % at
    // config 4: 768 items over 148 SMs); never more groups than vocabulary tiles
    static const int waves = std::max(1, env_int5("B200TRL_K5_WAVES", 16));
    int g = std::max(1, (waves * num_sms() + n_mtiles - 1) / n_mtiles);
    return std::min(g, n_ntiles);
}

}  // namespace
}  // namespace b200trl

using namespace b200trl;

extern "C" int64_t b200trl_fused_linear_workspace_bytes(int64_t n_rows, int64_t vocab) {
    if (n_rows <= 0 || vocab <= 0) return 0;
    const int64_t padded = ((n_rows + kBM - 1) / kBM) * kBM;
    const int64_t v1 = static_cast<int64_t>(groups_for(n_rows, vocab)) * padded * static_cast<int64_t>(sizeof(float4));
    return std::max(v1, tc_stats_workspace_bytes(n_rows, vocab));
}

extern "C" int b200trl_fused_linear_logprob_fwd(const void* hidden, int64_t hidden_row_stride, const void* weight,
                                                int64_t weight_row_stride, int64_t n_rows, int64_t hidden_size,
                                                int64_t vocab, const int64_t* ids, float inv_temperature,
                                                void* workspace, float* logp, float* entropy, float* lse,
                                                b200trl_stream_t stream) {
    B200TRL_REQUIRE(hidden && weight && ids && workspace && logp, B200TRL_E_INVALID, "fused_linear: null pointer");
    B200TRL_REQUIRE(n_rows > 0 && hidden_size > 0 && vocab > 0, B200TRL_E_INVALID, "fused_linear: bad shape");
    B200TRL_REQUIRE(hidden_size % 8 == 0 && hidden_row_stride % 8 == 0 && weight_row_stride % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(hidden) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0,
                    B200TRL_E_UNSUPPORTED, "fused_linear: bf16 operands need 16-byte aligned rows (hidden %% 8 == 0)");
    B200TRL_REQUIRE(inv_temperature > 0.f && std::isfinite(inv_temperature), B200TRL_E_INVALID,
                    "fused_linear: inv_temperature must be positive and finite");
    // The product path is the CTA-pair kernel (k7_tc_gemm.cu: cta_group::2, half the B tile per SM);
    // B200TRL_K5_IMPL=1 keeps the single-CTA kernel below for A/B measurements and as a cross-check in the tests.
    static const int impl = env_int5("B200TRL_K5_IMPL", 2);
    if (impl != 1) {
        const float c = static_cast<float>(static_cast<double>(inv_temperature) * 1.4426950408889634);
        int n_groups = 0;
        TcGemmParams p;
        p.epi = TC_EPI_STATS;
        p.A = hidden, p.lda = hidden_row_stride, p.B = weight, p.ldb = weight_row_stride;
        p.M = n_rows, p.N = vocab, p.K = hidden_size;
        p.ids = ids, p.c = c, p.partial = workspace, p.n_groups_out = &n_groups;
        int rc2 = tc_gemm(p, as_stream(stream));
        if (rc2) return rc2;
        return tc_merge_stats(workspace, n_groups, n_rows, c, logp, entropy, lse, as_stream(stream));
    }
    CUtensorMap map_a, map_b;
    int rc = make_map(&map_a, hidden, n_rows, hidden_size, hidden_row_stride, kBM);
    if (rc) return rc;
    rc = make_map(&map_b, weight, vocab, hidden_size, weight_row_stride, kBN);
    if (rc) return rc;
    K5Args a{};
    a.n_rows = n_rows;
    a.vocab = vocab;
    a.hidden = hidden_size;
    a.ids = ids;
    a.c = static_cast<float>(static_cast<double>(inv_temperature) * 1.4426950408889634);
    a.n_mtiles = static_cast<int>((n_rows + kBM - 1) / kBM);
    a.n_ntiles = static_cast<int>((vocab + kBN - 1) / kBN);
    a.n_groups = groups_for(n_rows, vocab);
    a.tiles_per_group = (a.n_ntiles + a.n_groups - 1) / a.n_groups;
    a.n_groups = (a.n_ntiles + a.tiles_per_group - 1) / a.tiles_per_group;  // drop empty trailing groups
    a.partial = static_cast<float4*>(workspace);
    const size_t smem = static_cast<size_t>(kStages) * kStageBytes + sizeof(Bars) + 1024;
    cudaError_t e = cudaFuncSetAttribute(k5_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
        set_error("fused_linear: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
        return B200TRL_E_LAUNCH;
    }
    const int grid = std::min(num_sms(), a.n_mtiles * a.n_groups);
    cudaStream_t s = as_stream(stream);
    k5_fwd_kernel<<<grid, kThreads, smem, s>>>(map_a, map_b, a);
    rc = check_launch("k5_fwd_kernel");
    if (rc) return rc;
    const int64_t padded = static_cast<int64_t>(a.n_mtiles) * kBM;
    k5_merge_kernel<<<static_cast<unsigned>((n_rows + 255) / 256), 256, 0, s>>>(a.partial, a.n_groups, padded, n_rows, a.c,
                                                                               logp, entropy, lse);
    return check_launch("k5_merge_kernel");
}
