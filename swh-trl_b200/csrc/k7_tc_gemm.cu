// K7: the CTA-pair tcgen05 GEMM family (BASELINE config 4: hidden 3584 -> V = 152064).
//
//   D[M, N] = A[M, K] * B[N, K]^T      bf16 operands, fp32 accumulation in tensor memory
//
// One kernel template, three epilogues, operands K-major or MN-major:
//   EPI_STATS  nothing is stored: every 128 x 256 accumulator tile is folded into the per-row online-softmax
//              statistics (m, S, U, selected logit) -> log-probs / entropies of hidden @ W^T with the logits never
//              leaving the SM (grpo_trainer.py:1163-1203 + :1258-1267 for the no-grad old / ref passes);
//   EPI_STORE  D is rounded to bf16 (+ bias) and stored row-major; optionally the same rounded values are folded into
//              the row statistics, i.e. exactly what a pass over the stored logits would compute
//              (the lm_head GEMM of the Liger seam, grpo_trainer.py:2005-2045; dH = dlogits W);
//   EPI_ACCUM  D is added into an fp32 matrix (dW += dlogits^T hidden across chunks of sequences) with TMA
//              reduce-add stores from swizzled 32 x 32 fp32 boxes: the read-modify-write happens in L2, the SM only
//              writes shared memory.  (A thread == row read-modify-write with ld/st.global touches 32 lines per warp
//              instruction: ncu showed the LSU data pipe 60 % busy and the tensor pipe down to 78 %.)
//   EPI_STORE32  the same boxes with plain TMA stores: D as fp32 (the first chunk of dW);
//   EPI_PARTIAL  split-K: the item's k-slice of D goes to its own fp32 plane (plain TMA store); a second small kernel
//              adds the planes in a fixed order and rounds to bf16 -> deterministic, and a contraction with few output
//              tiles and a huge K (dH: 224 tiles over 74 clusters, K = 152064) fills the last wave.
//
// Execution model (Blackwell-native, no legacy mma.sync anywhere):
//   * thread-block clusters of 2 CTAs on one TPC issue ONE tcgen05.mma.cta_group::2 per 256 x 256 x 16 step: each CTA
//     holds 128 rows of A and HALF of the B tile (128 of its 256 rows), so every SM stages 32 KB per 64-deep k-block
//     instead of 48 KB and reads 8 KB instead of 12 KB of shared memory per MMA;
//   * warp 0 (one lane) of each CTA: TMA producer, cp.async.bulk.tensor.2d.cta_group::2 into a 6-stage ring; the bytes
//     of BOTH CTAs complete the transaction barrier of the pair's leader;
//   * warp 1 (one lane) of the leader CTA: MMA issuer; tcgen05.commit ... multicast::cluster releases a stage in both
//     CTAs / publishes an accumulator to both epilogues;
//   * warp 2: allocates / frees the 512 TMEM columns (two 128 x 256 fp32 accumulators per CTA, ping-pong);
//   * warps 4-7: epilogue, tcgen05.ld 32 lanes x 32 columns, thread == accumulator row; overlaps the next tile's MMAs.
//     Stored tiles leave through shared memory: each warp writes its 32 rows into a private 128-byte-swizzled 4 KB box
//     (conflict-free) and one lane issues a TMA store / reduce-add, so global memory sees whole 128-byte rows
//     (direct st.global of the thread == row layout wrote 64-byte fragments and cost 5 % of the tensor time).
//   Persistent: cluster c takes work items c, c + n_clusters, ...; an item is (256-row block, run of n-tiles).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace b200trl {
namespace {

using namespace tc;

constexpr int kTileM = 128;  // accumulator rows per CTA (the pair covers 256)
constexpr int kTileN = 256;  // accumulator columns per tile
constexpr int kHalfN = 128;  // B rows staged by each CTA
constexpr int kTileK = 64;   // k-block: one 128-byte swizzle row of bf16
constexpr int kUmmaK = 16;
constexpr int kStages = 6;
constexpr int kOperandBytes = 128 * kTileK * 2;  // 16 KB: a 128 x 64 bf16 operand slab (either major)
constexpr int kStageBytes = 2 * kOperandBytes;   // per CTA
constexpr int kStagingBytes = 8 * 4096;          // epilogue (EPI_STORE): one 32-row x 128-byte box per warp
constexpr int kThreads = 384;  // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: two epilogue warp-groups
constexpr int kTmemCols = 512;
constexpr float kSlack = 6.0f;  // the running reference moves only when the tile maximum exceeds it by 2^6

enum { EPI_STATS = TC_EPI_STATS, EPI_STORE = TC_EPI_STORE, EPI_ACCUM = TC_EPI_ACCUM, EPI_PARTIAL = 3, EPI_STORE32 = 5 };

struct Bars {
    uint64_t full[kStages];   // leader CTA: both CTAs' TMA bytes of a stage have landed
    uint64_t empty[kStages];  // each CTA: the MMAs reading this stage have completed
    uint64_t tmem_full[2];    // each CTA: accumulator complete
    uint64_t tmem_empty[2];   // leader CTA: both epilogues have drained the accumulator (16 warp arrivals)
    uint32_t tmem_base;
};

struct GemmArgs {
    int64_t m_rows, n_cols, k_len;
    int n_mpairs, n_ntiles, n_groups, tiles_per_group, m_fastest;
    int sb_mpairs;  // 256-row blocks per super-block: the A rows that are walked together stay resident in L2
    int k_splits, kb_per_split;  // EPI_PARTIAL: item = (k-slice, row block, n-tile), slice-major so that all clusters
                                 // walk the same k-range at the same time (operands shared through L2)
    // statistics (EPI_STATS always; EPI_STORE when partial != nullptr)
    const int64_t* ids;
    float c;          // inv_T * log2(e)
    float4* partial;  // [n_groups * 2 column halves][n_mpairs * 256] : (m, S, U, selected logit or NaN)
    // output: EPI_STORE goes through the third tensor map (bf16 [m_rows, n_cols]); EPI_ACCUM: fp32 [m_rows, ldd], +=;
    // EPI_PARTIAL: fp32 planes [k_splits][m_rows, ldd], plane_stride elements apart
    void* out;
    int64_t ldd, plane_stride;
    const __nv_bfloat16* bias;  // EPI_STORE only, per column, may be null
    const float* addend;        // EPI_STORE only: fp32 [m_rows, ld_addend] added before the rounding, may be null
    int64_t ld_addend;          //   (the last chunk of dW: bf16(dW_fp32 + dlogits_c^T hidden_c) without a cast pass)
};

struct RowFold {
    float m, S, U, xsel;
};

// fold 32 logits of one row (columns cj .. cj + 31, `valid` of them real) into the row's running statistics
__device__ __forceinline__ void fold32(RowFold& f, const float* v, int valid, float c, int64_t id, int64_t cj) {
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (i < valid) ? v[i] : -INFINITY);
    mx *= c;
    if (mx > f.m + kSlack) {
        const float d = f.m - mx, s = ex2(d);
        f.U = s * fmaf(d, f.S, f.U);
        f.S *= s;
        f.m = mx;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float d = fmaf(v[i], c, -f.m);
        const float e = (i < valid) ? ex2(d) : 0.f;
        f.S += e;
        f.U = (i < valid) ? fmaf(e, d, f.U) : f.U;
    }
    if (id >= cj && id < cj + 32) {
        const int want = static_cast<int>(id - cj);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i == want) f.xsel = v[i];
    }
}

// Work order.  m_fastest: super-block of `sb_mpairs` row blocks -> run of n-tiles -> row block, so that clusters
// running at the same time share the B tile and re-read an A super-block that fits L2 (at config 4 the whole
// hidden matrix, 117 MB, does not: walking all of it per n-tile costs 30 % of the tensor throughput).  Otherwise
// row block -> run of n-tiles: clusters running at the same time share the A tile.
__device__ __host__ __forceinline__ void decode_item(int item, int n_mpairs, int n_groups, int sb_mpairs, int m_fastest,
                                                     int& g, int& mp) {
    if (m_fastest) {
        const int per_sb = sb_mpairs * n_groups;
        const int sb = item / per_sb, r = item - sb * per_sb;
        const int rest = n_mpairs - sb * sb_mpairs;
        const int mps = rest < sb_mpairs ? rest : sb_mpairs;
        g = r / mps;
        mp = sb * sb_mpairs + (r - g * mps);
    } else {
        mp = item / n_groups;
        g = item - mp * n_groups;
    }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <int kAMn, int kBMn, int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_d, const GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment in the shared window (1 KB of slack is allocated); both CTAs of the
    // pair compute the same offset, which the pair MMA relies on
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* staging = smem + kStages * kStageBytes;
    Bars& bars = *reinterpret_cast<Bars*>(staging + kStagingBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int items_per_split = a.n_mpairs * a.n_groups;
    const int n_items = items_per_split * a.k_splits;
    const int kblocks = static_cast<int>((a.k_len + kTileK - 1) / kTileK);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars.full[s], 1);
            mbar_init(&bars.empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars.tmem_full[b], 1);
            mbar_init(&bars.tmem_empty[b], 16);  // 8 epilogue warps x 2 CTAs
        }
        mbar_fence_init_cluster();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (kEpi == EPI_STORE || kEpi == EPI_ACCUM || kEpi == EPI_STORE32) tma_prefetch_desc(&map_d);
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_before_sync();
    cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / multicast commit
    fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars.tmem_base);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                int g, mp;
                const int ks = item / items_per_split;
                decode_item(item - ks * items_per_split, a.n_mpairs, a.n_groups, a.sb_mpairs, a.m_fastest, g, mp);
                const int kb0 = ks * a.kb_per_split, kb1 = min(kblocks, kb0 + a.kb_per_split);
                const int nt0 = g * a.tiles_per_group, nt1 = min(a.n_ntiles, nt0 + a.tiles_per_group);
                const int m0 = mp * (2 * kTileM) + static_cast<int>(rank) * kTileM;
                for (int nt = nt0; nt < nt1; ++nt) {
                    const int n0 = nt * kTileN + static_cast<int>(rank) * kHalfN;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&bars.empty[stage], phase ^ 1u);
                        unsigned char* sa = smem + stage * kStageBytes;
                        unsigned char* sb = sa + kOperandBytes;
                        const uint32_t full_leader = map_to_rank(smem_u32(&bars.full[stage]), 0);
                        if (leader) mbar_arrive_expect_tx(&bars.full[stage], 2 * kStageBytes);
                        if (kAMn) {  // rows = k, 64 m-elements per 128-byte row: two 64 x 64 boxes
                            tma_load_2d_pair(sa, &map_a, m0, kb * kTileK, full_leader);
                            tma_load_2d_pair(sa + kOperandBytes / 2, &map_a, m0 + 64, kb * kTileK, full_leader);
                        } else {
                            tma_load_2d_pair(sa, &map_a, kb * kTileK, m0, full_leader);
                        }
                        if (kBMn) {
                            tma_load_2d_pair(sb, &map_b, n0, kb * kTileK, full_leader);
                            tma_load_2d_pair(sb + kOperandBytes / 2, &map_b, n0 + 64, kb * kTileK, full_leader);
                        } else {
                            tma_load_2d_pair(sb, &map_b, kb * kTileK, n0, full_leader);
                        }
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
            // drain: every multicast commit aimed at this CTA's `empty` barriers has landed before the CTA may exit
            for (int s = 0; s < kStages; ++s) {
                mbar_wait(&bars.empty[stage], phase ^ 1u);
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA, one lane)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = instr_desc_bf16(2 * kTileM, kTileN, kAMn, kBMn);
            // K-major: 16 k-elements = 32 B inside the 128-byte swizzle row; MN-major: 16 k-rows = 2048 B
            constexpr uint64_t a_step = kAMn ? (16 * 128) >> 4 : 32 >> 4;
            constexpr uint64_t b_step = kBMn ? (16 * 128) >> 4 : 32 >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int buf = 0;
            uint32_t acc_phase = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                int g, mp;
                const int ks = item / items_per_split;
                decode_item(item - ks * items_per_split, a.n_mpairs, a.n_groups, a.sb_mpairs, a.m_fastest, g, mp);
                const int kb0 = ks * a.kb_per_split, kb1 = min(kblocks, kb0 + a.kb_per_split);
                const int nt0 = g * a.tiles_per_group, nt1 = min(a.n_ntiles, nt0 + a.tiles_per_group);
                for (int nt = nt0; nt < nt1; ++nt) {
                    mbar_wait_cluster(&bars.tmem_empty[buf], acc_phase ^ 1u);  // both epilogues drained this buffer
                    fence_after_sync();
                    const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * kTileN);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(&bars.full[stage], phase);
                        fence_after_sync();
                        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                        const uint64_t adesc = smem_desc_sw128(sa, kAMn ? kOperandBytes / 2 : 16);
                        const uint64_t bdesc = smem_desc_sw128(sa + kOperandBytes, kBMn ? kOperandBytes / 2 : 16);
#pragma unroll
                        for (int k = 0; k < kTileK / kUmmaK; ++k)
                            mma_pair_f16(tmem_d, adesc + a_step * k, bdesc + b_step * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
                        commit_pair_multicast(&bars.empty[stage], 3);  // both CTAs may refill the stage
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    commit_pair_multicast(&bars.tmem_full[buf], 3);  // accumulator complete -> both epilogues
                    buf ^= 1;
                    if (buf == 0) acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (both CTAs): thread == accumulator row
        // two warp-groups: warps 4-7 take columns [0, 128) of the tile, warps 8-11 columns [128, 256); a warp may
        // only read the TMEM lane quarter warp % 4
        const int ew = warp - 4;
        const int quarter = ew & 3, half = ew >> 2;
        const float c = a.c;
        const bool want_stats = (kEpi == EPI_STATS) || (kEpi == EPI_STORE && a.partial != nullptr);
        // this warp's staging box (32 rows x 128 bytes, SWIZZLE_128B: 16-byte chunk q of row r sits at q ^ (r & 7))
        unsigned char* my_box = staging + ew * 4096;
        const uint64_t stream_policy = l2_policy_evict_first();
        const uint32_t dst = smem_u32(my_box) + static_cast<uint32_t>(lane) * 128u;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        int buf = 0;
        uint32_t acc_phase = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            int g, mp;
            const int ks = item / items_per_split;
            decode_item(item - ks * items_per_split, a.n_mpairs, a.n_groups, a.sb_mpairs, a.m_fastest, g, mp);
            const int nt0 = g * a.tiles_per_group, nt1 = min(a.n_ntiles, nt0 + a.tiles_per_group);
            const int row0 = mp * (2 * kTileM) + static_cast<int>(rank) * kTileM + quarter * 32;  // this warp's first row
            const int64_t row = static_cast<int64_t>(row0) + lane;
            const bool row_ok = row < a.m_rows;
            const int64_t id = (want_stats && row_ok) ? a.ids[row] : -1;
            RowFold f{kNegBig, 0.f, 0.f, __int_as_float(0x7fc00000)};
            for (int nt = nt0; nt < nt1; ++nt) {
                const int col0 = nt * kTileN;
                float* out32 = nullptr;
                if (kEpi == EPI_PARTIAL) out32 = static_cast<float*>(a.out) + ks * a.plane_stride + row * a.ldd;
                mbar_wait(&bars.tmem_full[buf], acc_phase);
                fence_after_sync();
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * kTileN);
#pragma unroll 1
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = half * 4 + jj;
                    const int cj = col0 + j * 32;
                    const int64_t left = a.n_cols - cj;
                    const int valid = left < 32 ? static_cast<int>(left) : 32;  // columns past N are padding
                    if (valid <= 0) break;
                    float v[32];
                    tmem_ld32(tbase + static_cast<uint32_t>(j * 32), v);
                    if (kEpi == EPI_STATS) {
                        fold32(f, v, valid, c, id, cj);
                    } else if (kEpi == EPI_STORE) {
                        // a bf16 box is 64 columns: two TMEM loads fill its two halves
                        if ((jj & 1) == 0) {
                            if (lane == 0) bulk_wait_read<0>();  // the store that last read this box has finished
                            __syncwarp();
                        }
                        if (a.bias) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (i < valid) v[i] += __bfloat162float(a.bias[cj + i]);
                        }
                        if (a.addend && row_ok) {
                            const float* ad = a.addend + row * a.ld_addend + cj;
                            if (valid == 32) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    const float4 t = ld_stream_f4(ad + 4 * q);
                                    v[4 * q] += t.x, v[4 * q + 1] += t.y, v[4 * q + 2] += t.z, v[4 * q + 3] += t.w;
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (i < valid) v[i] += ad[i];
                            }
                        }
                        uint32_t p[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) p[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            st_shared_v4(dst + ((static_cast<uint32_t>((jj & 1) * 4 + q) ^ sw) << 4), p[4 * q], p[4 * q + 1],
                                         p[4 * q + 2], p[4 * q + 3]);
                        if (want_stats) {  // statistics of the ROUNDED logits: what a pass over the stored tile would see
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                v[2 * i] = __uint_as_float(p[i] << 16);
                                v[2 * i + 1] = __uint_as_float(p[i] & 0xffff0000u);
                            }
                            fold32(f, v, valid, c, id, cj);
                        }
                        if ((jj & 1) == 1 || cj + 32 >= a.n_cols) {  // box complete (or the row ends inside it)
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0 && row0 < a.m_rows) {
                                tma_store_2d_hint(&map_d, cj & ~63, row0, my_box, stream_policy);  // written once, read later
                                bulk_commit();
                            }
                        }
                    } else if (kEpi == EPI_ACCUM || kEpi == EPI_STORE32) {
                        // an fp32 box is 32 rows x 32 columns (128-byte rows); the TMA clips it at the matrix bounds
                        if (lane == 0) bulk_wait_read<0>();  // the copy that last read this box has finished
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            st_shared_v4(dst + ((static_cast<uint32_t>(q) ^ sw) << 4), __float_as_uint(v[4 * q]),
                                         __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                                         __float_as_uint(v[4 * q + 3]));
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && row0 < a.m_rows) {
                            if (kEpi == EPI_ACCUM)
                                tma_reduce_add_2d_hint(&map_d, cj, row0, my_box, stream_policy);
                            else
                                tma_store_2d_hint(&map_d, cj, row0, my_box, stream_policy);
                            bulk_commit();
                        }
                    } else {  // EPI_PARTIAL: this k-slice's plane, plain fp32 stores (a thread owns whole 128-byte lines)
                        if (row_ok) {
                            float* o = out32 + cj;
                            if (valid == 32) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) st_stream_f4(o + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                            } else {
                                for (int i = 0; i < valid; ++i) o[i] = v[i];
                            }
                        }
                    }
                }
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_rank(smem_u32(&bars.tmem_empty[buf]), 0));
                buf ^= 1;
                if (buf == 0) acc_phase ^= 1u;
            }
            if (want_stats && row_ok)
                a.partial[static_cast<int64_t>(g * 2 + half) * a.n_mpairs * (2 * kTileM) + row] = make_float4(f.m, f.S, f.U, f.xsel);
        }
        if ((kEpi == EPI_STORE || kEpi == EPI_ACCUM || kEpi == EPI_STORE32) && lane == 0)
            bulk_wait<0>();  // every store / reduction of this warp has reached global memory
    }
    fence_before_sync();
    cluster_sync_all();  // the peer's remote arrives and this CTA's multicast commits have all been consumed
    if (warp == 2) {
        fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

__global__ void tc_merge_kernel(const float4* __restrict__ partial, int n_groups, int64_t padded_rows, int64_t n_rows,
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

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// A bf16 matrix stored row-major as [outer, inner] with `ld` elements between rows; one box = 64 inner elements
// (128 bytes, SWIZZLE_128B) x `box_outer` rows.  K-major operand: inner = k, box_outer = 128 (m / n rows).
// MN-major operand: inner = m / n, box_outer = 64 (k rows).
int make_map(CUtensorMap* map, const void* base, int64_t outer, int64_t inner, int64_t ld, int box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("tc_gemm: cuTensorMapEncodeTiled is not available from the driver");
        return B200TRL_E_LAUNCH;
    }
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    const cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_outer)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("tc_gemm: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
        return B200TRL_E_INVALID;
    }
    return B200TRL_OK;
}

// Output map of EPI_STORE: bf16 [rows, cols], boxes of 32 rows x 64 columns (128 bytes), SWIZZLE_128B.
int make_out_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("tc_gemm: cuTensorMapEncodeTiled is not available from the driver");
        return B200TRL_E_LAUNCH;
    }
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    const cuuint32_t box[2] = {64u, 32u};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("tc_gemm: cuTensorMapEncodeTiled failed for the output (%d)", static_cast<int>(r));
        return B200TRL_E_INVALID;
    }
    return B200TRL_OK;
}

// Output map of EPI_ACCUM / EPI_STORE32: fp32 [rows, cols], boxes of 32 rows x 32 columns (128 bytes), SWIZZLE_128B.
int make_out_map_f32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("tc_gemm: cuTensorMapEncodeTiled is not available from the driver");
        return B200TRL_E_LAUNCH;
    }
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4};
    const cuuint32_t box[2] = {32u, 32u};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("tc_gemm: cuTensorMapEncodeTiled failed for the fp32 output (%d)", static_cast<int>(r));
        return B200TRL_E_INVALID;
    }
    return B200TRL_OK;
}

// bf16 out[r, c] = sum over the k-slices (in slice order) of the fp32 planes (+ bias[c]); 8 columns per thread
__global__ void __launch_bounds__(256) tc_splitk_finish_kernel(const float* __restrict__ ws, int splits, int64_t rows,
                                                               int64_t cols, __nv_bfloat16* __restrict__ out, int64_t ldd,
                                                               const __nv_bfloat16* __restrict__ bias) {
    const int64_t per_row = cols / 8;
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= rows * per_row) return;
    const int64_t r = idx / per_row, c8 = (idx - r * per_row) * 8;
    const int64_t plane = rows * cols;
    const float* p = ws + r * cols + c8;
    float4 lo = *reinterpret_cast<const float4*>(p), hi = *reinterpret_cast<const float4*>(p + 4);
    for (int s = 1; s < splits; ++s) {
        const float4 a = *reinterpret_cast<const float4*>(p + s * plane), b = *reinterpret_cast<const float4*>(p + s * plane + 4);
        lo.x += a.x, lo.y += a.y, lo.z += a.z, lo.w += a.w;
        hi.x += b.x, hi.y += b.y, hi.z += b.z, hi.w += b.w;
    }
    if (bias) {
        const uint4 bb = *reinterpret_cast<const uint4*>(bias + c8);
        lo.x += __uint_as_float(bb.x << 16), lo.y += __uint_as_float(bb.x & 0xffff0000u);
        lo.z += __uint_as_float(bb.y << 16), lo.w += __uint_as_float(bb.y & 0xffff0000u);
        hi.x += __uint_as_float(bb.z << 16), hi.y += __uint_as_float(bb.z & 0xffff0000u);
        hi.z += __uint_as_float(bb.w << 16), hi.w += __uint_as_float(bb.w & 0xffff0000u);
    }
    *reinterpret_cast<uint4*>(out + r * ldd + c8) =
        make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

int n_clusters_for_device() { return std::max(1, num_sms() / 2); }

// Static round-robin of (256-row block, run of n-tiles) items over the clusters: pick the run length whose slowest
// cluster finishes first (an item costs its tiles plus ~0.15 tile of pipeline restart and partial write).
int plan_tiles_per_group(int n_mpairs, int n_ntiles, int n_clusters, int sb_mpairs) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int>, int> cache;
    const auto key = std::make_tuple(n_mpairs, n_ntiles, n_clusters, sb_mpairs);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
    }
    const int forced = env_int("B200TRL_K7_TPG", 0);
    const int lo = std::min(n_ntiles, 4);
    int best = std::max(lo, std::min(n_ntiles, forced));
    if (forced <= 0) {
        double best_cost = 1e300;
        for (int tpg = lo; tpg <= n_ntiles; ++tpg) {
            const int groups = (n_ntiles + tpg - 1) / tpg;
            const int last = n_ntiles - (groups - 1) * tpg;
            const int64_t items = static_cast<int64_t>(groups) * n_mpairs;
            if (items > 65536) continue;
            double worst = 0;
            for (int cl = 0; cl < std::min<int64_t>(n_clusters, items); ++cl) {
                double t = 0;
                for (int64_t i = cl; i < items; i += n_clusters) {
                    int g, mp;
                    decode_item(static_cast<int>(i), n_mpairs, groups, sb_mpairs, 1, g, mp);
                    t += ((g == groups - 1) ? last : tpg) + 0.15;
                }
                worst = std::max(worst, t);
            }
            if (worst < best_cost - 1e-9) {
                best_cost = worst;
                best = tpg;
            }
        }
    }
    std::lock_guard<std::mutex> lock(mu);
    cache[key] = best;
    return best;
}

template <int kAMn, int kBMn, int kEpi>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const GemmArgs& a, cudaStream_t s) {
    const size_t smem = static_cast<size_t>(kStages) * kStageBytes + kStagingBytes + sizeof(Bars) + 1024;
    auto kern = tc_gemm_kernel<kAMn, kBMn, kEpi>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
        set_error("tc_gemm: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
        return B200TRL_E_LAUNCH;
    }
    const int64_t items = static_cast<int64_t>(a.n_mpairs) * a.n_groups * a.k_splits;
    const int clusters = static_cast<int>(std::min<int64_t>(n_clusters_for_device(), items));
    kern<<<2 * clusters, kThreads, smem, s>>>(ma, mb, md, a);
    return check_launch("tc_gemm_kernel");
}

// Split-K decision for a stored GEMM: with `tiles` output tiles over `clusters` clusters the last wave of whole tiles
// may be nearly empty (dH at config 4: 224 tiles on 74 clusters = 3.03 waves, paid as 4).  Splitting K into S slices
// gives S * tiles items of 1/S the length; the planes cost S * M * N * 8 bytes of extra traffic.  Returns S (1 = none).
int plan_k_splits(int64_t tiles, int clusters, int kblocks, int64_t M, int64_t N, int64_t ws_bytes) {
    static const int forced = env_int("B200TRL_K7_SPLITK", 0);
    const double tile_s = 2.0 * 256 * 256 * kblocks * kTileK / (2 * 10.5e12);  // one tile on one CTA pair
    int best = 1;
    double best_t = static_cast<double>((tiles + clusters - 1) / clusters) * tile_s;
    for (int S = 2; S <= 16; ++S) {
        if (kblocks / S < 32 || S * M * N * 4 > ws_bytes) break;
        const int kbs = (kblocks + S - 1) / S;
        const int S_eff = (kblocks + kbs - 1) / kbs;
        const double rounds = static_cast<double>((tiles * S_eff + clusters - 1) / clusters);
        const double t = rounds * tile_s * kbs / kblocks + S_eff * M * N * 8.0 / 6.0e12 + 4e-6;
        if ((forced == S) || (forced <= 0 && t < best_t * 0.985)) {
            best_t = t;
            best = S;
            if (forced == S) break;
        }
    }
    return best;
}

}  // namespace

// D = A B^T with the epilogue of `p.epi`; shared by the C entry points of this file, k5 and k6 (the seam).
int tc_gemm(const TcGemmParams& p, cudaStream_t s) {
    const int64_t M = p.M, N = p.N, K = p.K;
    B200TRL_REQUIRE(M > 0 && N > 0 && K > 0, B200TRL_E_INVALID, "tc_gemm: bad shape");
    B200TRL_REQUIRE(p.lda % 8 == 0 && p.ldb % 8 == 0 && (reinterpret_cast<uintptr_t>(p.A) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(p.B) & 15) == 0,
                    B200TRL_E_UNSUPPORTED, "tc_gemm: bf16 operands need 16-byte aligned rows");
    B200TRL_REQUIRE(p.epi != TC_EPI_STATS || (p.partial && p.ids), B200TRL_E_INVALID,
                    "tc_gemm: statistics need ids and a workspace");
    const bool f32_out = p.epi == TC_EPI_ACCUM || p.epi == TC_EPI_STORE_F32;
    B200TRL_REQUIRE(p.epi == TC_EPI_STATS ||
                        (p.out && p.ldd % (f32_out ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0),
                    B200TRL_E_UNSUPPORTED, "tc_gemm: output rows must be 16-byte aligned");
    B200TRL_REQUIRE(!p.addend || (p.epi == TC_EPI_STORE && p.ld_addend % 4 == 0 &&
                                  (reinterpret_cast<uintptr_t>(p.addend) & 15) == 0),
                    B200TRL_E_UNSUPPORTED, "tc_gemm: the fp32 addend needs the bf16 output and 16-byte aligned rows");
    CUtensorMap ma, mb;
    int rc = p.a_mn ? make_map(&ma, p.A, K, M, p.lda, 64) : make_map(&ma, p.A, M, K, p.lda, 128);
    if (rc) return rc;
    rc = p.b_mn ? make_map(&mb, p.B, K, N, p.ldb, 64) : make_map(&mb, p.B, N, K, p.ldb, 128);
    if (rc) return rc;
    GemmArgs a{};
    a.m_rows = M;
    a.n_cols = N;
    a.k_len = K;
    a.n_mpairs = static_cast<int>((M + 2 * kTileM - 1) / (2 * kTileM));
    a.n_ntiles = static_cast<int>((N + kTileN - 1) / kTileN);
    // A super-block = the 256-row blocks whose A rows (sb_mpairs * 256 * K bf16) fit a 32 MB share of L2, evened out
    {
        static const int budget_mb = std::max(1, env_int("B200TRL_K7_SB_MB", 32));
        const int64_t per_mpair = 2 * kTileM * K * 2;
        const int fit = static_cast<int>(std::max<int64_t>(1, (static_cast<int64_t>(budget_mb) << 20) / per_mpair));
        const int n_sb = (a.n_mpairs + fit - 1) / fit;
        a.sb_mpairs = (a.n_mpairs + n_sb - 1) / n_sb;
    }
    const bool stats = p.partial != nullptr;
    if (stats) {
        a.tiles_per_group = plan_tiles_per_group(a.n_mpairs, a.n_ntiles, n_clusters_for_device(), a.sb_mpairs);
        a.m_fastest = 1;
    } else {
        a.tiles_per_group = 1;
        a.m_fastest = p.m_fastest;
    }
    a.n_groups = (a.n_ntiles + a.tiles_per_group - 1) / a.tiles_per_group;
    a.ids = p.ids;
    a.c = p.c;
    a.partial = static_cast<float4*>(p.partial);
    a.bias = static_cast<const __nv_bfloat16*>(p.bias);
    a.addend = p.addend;
    a.ld_addend = p.ld_addend;
    a.k_splits = 1;
    a.kb_per_split = static_cast<int>((K + kTileK - 1) / kTileK);
    if (p.n_groups_out) *p.n_groups_out = a.n_groups * 2;  // one partial per (run of n-tiles, column half)
    CUtensorMap md = ma;  // only EPI_STORE stores through a tensor map
    if (p.epi == TC_EPI_STORE && p.splitk_ws && !stats && !p.addend && N % 8 == 0 && p.a_mn == 0) {
        const int kblocks = a.kb_per_split;
        const int S = plan_k_splits(static_cast<int64_t>(a.n_mpairs) * a.n_ntiles, n_clusters_for_device(), kblocks, M, N,
                                    p.splitk_ws_bytes);
        if (S > 1) {
            a.kb_per_split = (kblocks + S - 1) / S;
            a.k_splits = (kblocks + a.kb_per_split - 1) / a.kb_per_split;
            a.bias = nullptr;  // added once, by the finishing kernel
            a.out = p.splitk_ws;
            a.ldd = N;
            a.plane_stride = M * N;
            rc = p.b_mn ? launch<0, 1, EPI_PARTIAL>(ma, mb, md, a, s) : launch<0, 0, EPI_PARTIAL>(ma, mb, md, a, s);
            if (rc) return rc;
            const int64_t n_thr = M * (N / 8);
            tc_splitk_finish_kernel<<<static_cast<unsigned>((n_thr + 255) / 256), 256, 0, s>>>(
                static_cast<const float*>(p.splitk_ws), a.k_splits, M, N, static_cast<__nv_bfloat16*>(p.out), p.ldd,
                static_cast<const __nv_bfloat16*>(p.bias));
            return check_launch("tc_splitk_finish_kernel");
        }
    }
    if (p.epi == TC_EPI_STORE) rc = make_out_map(&md, p.out, M, N, p.ldd);
    if (p.epi == TC_EPI_ACCUM || p.epi == TC_EPI_STORE_F32) rc = make_out_map_f32(&md, p.out, M, N, p.ldd);
    if (rc) return rc;
    a.out = p.out;
    a.ldd = p.ldd;
    a.plane_stride = 0;
    const int key = p.a_mn * 100 + p.b_mn * 10 + p.epi;
    switch (key) {
        case 0 * 100 + 0 * 10 + TC_EPI_STATS: return launch<0, 0, EPI_STATS>(ma, mb, md, a, s);
        case 0 * 100 + 0 * 10 + TC_EPI_STORE: return launch<0, 0, EPI_STORE>(ma, mb, md, a, s);
        case 0 * 100 + 1 * 10 + TC_EPI_STORE: return launch<0, 1, EPI_STORE>(ma, mb, md, a, s);
        case 1 * 100 + 1 * 10 + TC_EPI_STORE: return launch<1, 1, EPI_STORE>(ma, mb, md, a, s);
        case 1 * 100 + 1 * 10 + TC_EPI_ACCUM: return launch<1, 1, EPI_ACCUM>(ma, mb, md, a, s);
        case 1 * 100 + 1 * 10 + TC_EPI_STORE_F32: return launch<1, 1, EPI_STORE32>(ma, mb, md, a, s);
        default: break;
    }
    set_error("tc_gemm: unsupported operand layout / epilogue combination (%d, %d, %d)", p.a_mn, p.b_mn, p.epi);
    return B200TRL_E_UNSUPPORTED;
}

int64_t tc_stats_workspace_bytes(int64_t n_rows, int64_t n_cols) {
    if (n_rows <= 0 || n_cols <= 0) return 0;
    const int64_t padded = ((n_rows + 2 * kTileM - 1) / (2 * kTileM)) * (2 * kTileM);
    const int64_t n_ntiles = (n_cols + kTileN - 1) / kTileN;
    const int64_t max_groups = (n_ntiles + 3) / 4 + 1;  // plan_tiles_per_group never goes below runs of 4 (or all)
    return 2 * max_groups * padded * static_cast<int64_t>(sizeof(float4));  // two column halves per run
}

int tc_merge_stats(const void* partial, int n_groups, int64_t n_rows, float c, float* logp, float* entropy, float* lse,
                   cudaStream_t s) {
    const int64_t padded = ((n_rows + 2 * kTileM - 1) / (2 * kTileM)) * (2 * kTileM);
    tc_merge_kernel<<<static_cast<unsigned>((n_rows + 255) / 256), 256, 0, s>>>(static_cast<const float4*>(partial), n_groups,
                                                                               padded, n_rows, c, logp, entropy, lse);
    return check_launch("tc_merge_kernel");
}

}  // namespace b200trl

using namespace b200trl;

extern "C" int64_t b200trl_tc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int out_kind) {
    if (M <= 0 || N <= 0 || K <= 0 || out_kind != B200TRL_TC_OUT_BF16) return 0;
    // room for up to 8 fp32 planes of D, capped at 512 MB: enough for every split the planner would pick
    return std::min<int64_t>(8 * M * N * 4, int64_t(512) << 20);
}

extern "C" int b200trl_tc_gemm(const void* A, int a_layout, int64_t lda, const void* B, int b_layout, int64_t ldb,
                               int64_t M, int64_t N, int64_t K, int out_kind, void* out, int64_t ldd, const void* bias,
                               const float* addend, int64_t ld_addend, int m_fastest, void* workspace,
                               int64_t workspace_bytes, b200trl_stream_t stream) {
    B200TRL_REQUIRE(A && B && out, B200TRL_E_INVALID, "tc_gemm: null pointer");
    B200TRL_REQUIRE(out_kind == B200TRL_TC_OUT_BF16 || out_kind == B200TRL_TC_OUT_F32_ACC || out_kind == B200TRL_TC_OUT_F32,
                    B200TRL_E_INVALID, "tc_gemm: unknown out_kind %d", out_kind);
    B200TRL_REQUIRE((!bias && !addend) || out_kind == B200TRL_TC_OUT_BF16, B200TRL_E_INVALID,
                    "tc_gemm: bias / addend need the bf16 output");
    TcGemmParams p;
    p.a_mn = a_layout;
    p.b_mn = b_layout;
    p.epi = out_kind == B200TRL_TC_OUT_BF16 ? TC_EPI_STORE : out_kind == B200TRL_TC_OUT_F32_ACC ? TC_EPI_ACCUM : TC_EPI_STORE_F32;
    p.A = A, p.lda = lda, p.B = B, p.ldb = ldb, p.M = M, p.N = N, p.K = K;
    p.out = out, p.ldd = ldd, p.bias = bias, p.addend = addend, p.ld_addend = ld_addend;
    p.m_fastest = m_fastest;
    p.splitk_ws = workspace;
    p.splitk_ws_bytes = workspace ? workspace_bytes : 0;
    return tc_gemm(p, as_stream(stream));
}
