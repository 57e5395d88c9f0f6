// K1 "resident" implementation (the product path for bf16 logits): every logit row is read from HBM
// exactly once and dlogits written exactly once.
//
//   * each CTA pulls its share of a row with 16-24 KB TMA bulk copies (cp.async.bulk -> UBLKCP) into a
//     ring of shared-memory slots, signalled through mbarriers; one DMA warp (one elected lane) issues
//     every bulk copy;
//   * 8-24 consumer warps fold the chunks into an online (reference, sum, sum*delta) partial as they
//     land (packed bf16 max, f32x2 FMA/ADD, MUFU ex2); a reducer warp finishes the row: log-prob,
//     entropy, lse and the per-token d(loss)/d(logp) (GRPO or PPO surrogate, see k1_args.cuh);
//   * FUSED pass (forward + dlogits in one read): the row must stay on chip between the two sweeps,
//     so it is split over a thread-block cluster (1/2/4/8 CTAs, picked so a slice fits one CTA's
//     shared memory); the CTAs swap their partials with one st.async DSMEM message per peer, and the
//     consumers turn the still-resident slice into dlogits that leave straight from registers
//     (st.global.cs), releasing each slot for the next row's chunk as soon as it has been read;
//   * FORWARD-ONLY and BACKWARD-ONLY passes keep nothing resident: a row streams through the ring of
//     a single CTA (no cluster, any vocabulary); backward-only rewrites chunks in place and stores
//     them with TMA bulk stores.
//
// Persistent grid: clusters (or single CTAs) loop over rows with a static stride.  Replaces
// trl/trainer/utils.py:1430-1490 + grpo_trainer.py:1258 + the autograd backward down to the logits
// (see include/b200trl.h).  Measured behaviour, the phase trace and what was tried: DESIGN.md §3.
// Two translation units are built from this file so that the ~170 kernel instantiations compile in parallel:
// k1_resident.cu itself (bf16, K1_UNIT_F16 = 0: every variant plus the host-side policy and the public entry points)
// and k1_resident_f16.cu (defines K1_UNIT_F16 = 1 and includes this file: the fp16 instantiations only).
#ifndef K1_UNIT_F16
#define K1_UNIT_F16 0
#endif
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "k1_args.cuh"

namespace b200trl {
namespace {

constexpr int kMaxConsumers = 768;  // consumer threads per CTA: 768 / 640 / 512 (1 CTA / SM), 256 (2 CTAs / SM)
constexpr int kChunkBytes = 16384;  // default chunk; per-geometry value: chunk_bytes_for()
constexpr int kMaxSlots = 13;
__host__ __device__ constexpr int chunk_bytes_for(int consumers) {
    // 320 / 384 consumers ("lean" shapes): the 640 / 768 chunk with FOUR vectors per thread instead of two
    return (consumers == 768 || consumers == 640) ? consumers * 32
                                                  : ((consumers == 320 || consumers == 384) ? consumers * 64 : 16384);
}
constexpr int kMaxCluster = 8;
constexpr int kMaxCountSeqs = 256;  // in-kernel mask statistics (K1Args::count_mask): sequences per batch
constexpr float kSlack = 6.0f;   // reference point may trail the running max by 2^6

struct __align__(16) Part4 {
    float m, s, u, pad;
};

// What the consumers need to turn a resident slice into dlogits (written by the reducer warp).
struct __align__(16) RowResult {
    float lse2;    // log2-domain logsumexp of the scaled row
    float ng;      // -g' = -(d loss / d logp) * inv_T   (0 => the row's dlogits are zero)
    float patch;   // value at the selected id: g' * (1 - p_id)
    int id_chunk;  // chunk of this CTA's slice holding the selected id, or -1
    int id_vec;    // 16-byte vector inside that chunk
    int id_elem;   // element inside that chunk
    int pad0, pad1;
};

// ------------------------------------------------------------------ diagnostics
// Phase timestamps (b200trl_k1_set_trace), compiled in only with -DB200TRL_K1_TRACE (`make trace` ->
// lib/libb200trl_trace.so; the hooks cost the production kernel ~20 % even when switched off, so it has none).
// Event = tag << 56 | row << 40 | chunk << 32 | low 32 bits of clock64, kept in a small shared-memory log per role
// (0 consumer warp 0, 1 reducer, 2 DMA) and copied out when the CTA finishes; first four CTAs, kTraceRows rows
// starting at trace_row0.
constexpr int kTraceCtas = 4, kTraceRoles = 3, kTraceRows = 8, kTraceEvents = 168;
#ifdef B200TRL_K1_TRACE
struct TraceLog {
    unsigned long long ev[kTraceRoles][kTraceEvents];
    int n[kTraceRoles];
};
struct Tracer {
    TraceLog* log;
    int role, n, row0;
    __device__ __forceinline__ void init(const K1Args& a, TraceLog* l, int r, bool active) {
        log = (a.trace && active && blockIdx.x < kTraceCtas) ? l : nullptr;
        role = r;
        n = 0;
        row0 = a.trace_row0;
    }
    __device__ __forceinline__ void ev(int tag, int row, int chunk) {
        if (log && n < kTraceEvents && row >= row0 && row < row0 + kTraceRows) {
            const unsigned long long t = static_cast<unsigned long long>(clock64()) & 0xffffffffull;
            log->ev[role][n++] = (static_cast<unsigned long long>(tag) << 56) |
                                 (static_cast<unsigned long long>(row & 0xffff) << 40) |
                                 (static_cast<unsigned long long>(chunk & 0xff) << 32) | t;
        }
    }
    __device__ __forceinline__ void finish() {
        if (log) log->n[role] = n;
    }
};
#else
struct TraceLog {};
struct Tracer {
    __device__ __forceinline__ void init(const K1Args&, TraceLog*, int, bool) {}
    __device__ __forceinline__ void ev(int, int, int) {}
    __device__ __forceinline__ void finish() {}
};
#endif

struct Smem {
    // slots first (16 KB each, 128-byte aligned)
    uint64_t full_bar[kMaxSlots];  // DMA -> consumers: chunk landed (tx-count)
    uint64_t done_bar[kMaxSlots];  // consumers -> DMA: slot may be stored / refilled (one arrive per warp)
    uint64_t part_bar[2];          // consumers -> reducer: 16 warp partials of a row are in warp_part
    uint64_t res_bar[2];           // reducer -> consumers: RowResult of a row is ready
    uint64_t xchg_bar[2];          // peers -> reducer: every CTA of the cluster delivered its partial
    Part4 xchg[2][kMaxCluster];
    Part4 warp_part[2][kMaxConsumers / 32];
    RowResult result[2];
    float ppo_count;
    float ppo_count_p1;            // G_PPO step sums: positions with t <= len + 1 (ppo_trainer.py:505)
    float total_cnt;               // count_mask: number of unmasked tokens of the batch ...
    float row_cnt[kMaxCountSeqs];  // ... and per sequence, counted by the consumer warps while the first chunks fly
    TraceLog trace;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// B200TRL_K1_WAIT_HINT (ns): suspend-time hint of mbarrier.try_wait -- the warp sleeps in hardware until the phase
// completes (or the hint expires) instead of re-issuing the probe every ~30 cycles.
#ifndef B200TRL_K1_WAIT_HINT
#define B200TRL_K1_WAIT_HINT 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if B200TRL_K1_WAIT_HINT > 0
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(B200TRL_K1_WAIT_HINT)
        : "memory");
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
#endif
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// One-way DSMEM message: 16 bytes land in a peer's shared memory and complete 16 bytes of its mbarrier's transaction
// count -- no release/acquire round trip on the sender's side.
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     cluster_addr),
                 "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_bar)
                 : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
template <int NC>
__device__ __forceinline__ void consumer_bar() {
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
}

// packed 2 x fp32 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// ------------------------------------------------------------------ element type (bf16, or fp16 when F16)
// Only four things depend on the 16-bit element type: the packed maximum, the unpack to fp32 pairs, the pack of a
// gradient pair, and the "very negative" stand-in of a masked element.  fp16 runs on the generic consumer code.
template <bool F16>
__device__ __forceinline__ uint32_t el_max2(uint32_t a, uint32_t b) {
    uint32_t d;
    if (F16)
        asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else
        asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
template <bool F16>
__device__ __forceinline__ uint64_t el_unpack(uint32_t w) {  // packed pair -> (lo, hi) as f32x2
    if (F16) {
        float lo, hi;
        asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}"
            : "=f"(lo), "=f"(hi)
            : "r"(w));
        return pack2(lo, hi);
    }
    return pack2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <bool F16>
__device__ __forceinline__ float el_pair_max(uint32_t mx) {  // max of the two halves of a packed pair, as fp32
    float lo, hi;
    unpack2(el_unpack<F16>(mx), lo, hi);
    return fmaxf(lo, hi);
}
template <bool F16>
__device__ __forceinline__ uint32_t el_pack(float lo, float hi) {
    uint32_t d;
    if (F16)
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
template <bool F16>
__device__ __forceinline__ void el_store(void* base, int64_t idx, float v) {
    if (F16)
        reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
    else
        reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
}
// stand-in of a masked element: bf16(-1e30) / the most negative finite fp16 (-65504): 2^(x c - m) == 0 and
// 0 * (x c - m) == -0 for any sane inv_T
template <bool F16>
constexpr uint32_t kNegEl = F16 ? 0xFBFFu : 0xF149u;

// ------------------------------------------------------------------ consumer state
struct Acc {
    float m;       // reference point (log2 units)
    uint64_t s2;   // running sums of 2^(y-m)        (two independent chains: s2/u2 and t2/v2)
    uint64_t u2;   // running sums of 2^(y-m)*(y-m)
    uint64_t t2;
    uint64_t v2;
};

__device__ __forceinline__ Acc acc_empty() {
    return Acc{kNegBig, 0ull, 0ull, 0ull, 0ull};  // bit pattern 0 == (0.f, 0.f)
}

template <bool F16 = false>
__device__ __forceinline__ uint32_t vec_max(const uint4& v) {
    return el_max2<F16>(el_max2<F16>(v.x, v.y), el_max2<F16>(v.z, v.w));
}

// move the reference point to cm (rare after the first chunk)
__device__ __forceinline__ void acc_rescale(Acc& a, float cm) {
    const float d = a.m - cm;
    const float f = ex2(d);
    const uint64_t f2 = pack2(f, f), d2 = pack2(d, d);
    a.u2 = fmul2(f2, ffma2(d2, a.s2, a.u2));
    a.s2 = fmul2(a.s2, f2);
    a.v2 = fmul2(f2, ffma2(d2, a.t2, a.v2));
    a.t2 = fmul2(a.t2, f2);
    a.m = cm;
}

// ENT = false: the caller wants no entropy (the no-grad old / ref log-prob passes), so the sum of 2^(y-m) (y-m) is not
// accumulated -- one packed FMA per pair less in a pass whose speed under the power cap is set by the energy it spends
template <bool F16 = false, bool ENT = true>
__device__ __forceinline__ void acc_words(uint64_t& s2, uint64_t& u2, const uint4& v, uint64_t c2, uint64_t nm2) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t x2 = el_unpack<F16>(w[i]);
        const uint64_t d2 = ffma2(x2, c2, nm2);
        float d0, d1;
        unpack2(d2, d0, d1);
        const uint64_t e2 = pack2(ex2(d0), ex2(d1));
        s2 = fadd2(s2, e2);
        if (ENT) u2 = ffma2(e2, d2, u2);
    }
}

// fold one 16-byte vector
template <bool F16 = false, bool ENT = true>
__device__ __forceinline__ void acc_vec(Acc& a, const uint4& v, float c, uint64_t c2) {
    const float cm = el_pair_max<F16>(vec_max<F16>(v)) * c;
    if (cm > a.m + kSlack) acc_rescale(a, cm);
    acc_words<F16, ENT>(a.s2, a.u2, v, c2, pack2(-a.m, -a.m));
}

// fold two vectors with ONE reference-point check and two independent accumulation chains
template <bool F16 = false, bool ENT = true>
__device__ __forceinline__ void acc_vec2(Acc& a, const uint4& v0, const uint4& v1, float c, uint64_t c2) {
    const float cm = el_pair_max<F16>(el_max2<F16>(vec_max<F16>(v0), vec_max<F16>(v1))) * c;
    if (cm > a.m + kSlack) acc_rescale(a, cm);
    const uint64_t nm2 = pack2(-a.m, -a.m);
    acc_words<F16, ENT>(a.s2, a.u2, v0, c2, nm2);
    acc_words<F16, ENT>(a.t2, a.v2, v1, c2, nm2);
}

template <bool F16 = false>
__device__ __forceinline__ uint4 grad_vec(const uint4& v, uint64_t c2, uint64_t nl2, uint64_t ng2) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t x2 = el_unpack<F16>(w[i]);
        const uint64_t d2 = ffma2(x2, c2, nl2);
        float d0, d1;
        unpack2(d2, d0, d1);
        const uint64_t e2 = pack2(ex2(d0), ex2(d1));
        float o0, o1;
        unpack2(fmul2(e2, ng2), o0, o1);
        o[i] = el_pack<F16>(o0, o1);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}
// ---- skewed rows (SKEW): the row does not start on a 16-byte boundary (vocab % 8 != 0, e.g. GPT-2's 50 257, or a
// strided view).  The bulk copies fetch the 16-byte-aligned span that CONTAINS the slice, so the slice starts `h`
// elements (0..7) into the first vector and the last vector may end early; the elements outside the slice belong to
// the neighbouring rows (or lie past the tensor, inside the same 16-byte granule: never a fault) and are replaced by a
// large negative logit before any arithmetic.  Only the first and the last vector of a slice are affected.
template <bool F16 = false>
__device__ __forceinline__ uint4 mask_vec(const uint4& v, int lo, int hi) {  // keep elements [lo, hi) of the 8
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (2 * i < lo || 2 * i >= hi) w[i] = (w[i] & 0xffff0000u) | kNegEl<F16>;
        if (2 * i + 1 < lo || 2 * i + 1 >= hi) w[i] = (w[i] & 0x0000ffffu) | (kNegEl<F16> << 16);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
// elements [lo, hi) of one 16-byte vector, as 2-byte stores (the rest of the granule is not ours to write)
__device__ __forceinline__ void st_global_edge(uint4* p, const uint4& v, int lo, int hi) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    unsigned short* q = reinterpret_cast<unsigned short*>(p);
#pragma unroll
    for (int e = 0; e < 8; ++e)
        if (e >= lo && e < hi) q[e] = static_cast<unsigned short>((e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu));
}
// streaming 16-byte store: dlogits are not read again by this kernel
__device__ __forceinline__ void st_global_cs(uint4* p, const uint4& v) {
    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ Partial acc_to_partial(const Acc& a) {
    float s0, s1, u0, u1;
    unpack2(fadd2(a.s2, a.t2), s0, s1);
    unpack2(fadd2(a.u2, a.v2), u0, u1);
    return Partial{a.m, s0 + s1, u0 + u1};
}

// ------------------------------------------------------------------ fast consumer of the fused pass
// The product path (fused forward + dlogits, direct stores, two accumulation chains) has its own consumer code:
//   * shared memory is addressed with 32-bit shared-window addresses computed once (the generic-pointer form made
//     the compiler rebuild the window base for every chunk) and the full chunks run in loops that know nothing about
//     the partial last chunk: ~40 -> ~12 bookkeeping instructions per chunk and warp, a quarter of all instructions;
//   * one reference-point check per chunk, driven by a PACKED running maximum (bf16x2): the common case is four
//     3-input packed max instructions and one integer compare;
//   Tried and dropped: writing bf16(2^(y - m)) back over the folded logits so that the backward sweep is a multiply
//   instead of a second exponential (half the MUFU load).  The cached value is rounded (2^-8) before the gradient is
//   rounded (2^-8): 28 of 884 736 elements of test_ppo_fused_large_vocab left the one-ulp bar (1.16 % vs 0.78 %), and
//   the pass was 10 % SLOWER (1.86 vs 1.68 ms burst: the write-back competes with the TMA for shared-memory bandwidth).
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
#if B200TRL_K1_WAIT_HINT > 0
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_U:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE_U;\n\t"
        "bra WAIT_LOOP_U;\n\t"
        "WAIT_DONE_U:\n\t"
        "}" ::"r"(bar),
        "r"(parity), "r"(B200TRL_K1_WAIT_HINT)
        : "memory");
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_U:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE_U;\n\t"
        "bra WAIT_LOOP_U;\n\t"
        "WAIT_DONE_U:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr uint32_t kNegInf2 = 0xff80ff80u;  // bf16x2 (-inf, -inf)

// fold one vector
#ifndef B200TRL_K1_TREE
#define B200TRL_K1_TREE 0
#endif
__device__ __forceinline__ void fold_words(uint64_t& s2, uint64_t& u2, const uint4& v, uint64_t c2, uint64_t nm2) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#if B200TRL_K1_TREE
    // A/B variant: the four pairs of a vector are summed as a tree, so the loop-carried accumulators see ONE dependent
    // add / fma per vector instead of four
    uint64_t e2[4], d2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t x2 = pack2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
        d2[i] = ffma2(x2, c2, nm2);
        float d0, d1;
        unpack2(d2[i], d0, d1);
        e2[i] = pack2(ex2(d0), ex2(d1));
    }
    s2 = fadd2(s2, fadd2(fadd2(e2[0], e2[1]), fadd2(e2[2], e2[3])));
    const uint64_t ua = ffma2(e2[1], d2[1], fmul2(e2[0], d2[0]));
    const uint64_t ub = ffma2(e2[3], d2[3], fmul2(e2[2], d2[2]));
    u2 = fadd2(u2, fadd2(ua, ub));
#else
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t x2 = pack2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
        const uint64_t d2 = ffma2(x2, c2, nm2);
        float d0, d1;
        unpack2(d2, d0, d1);
        const float e0 = ex2(d0), e1 = ex2(d1);
        const uint64_t e2 = pack2(e0, e1);
        s2 = fadd2(s2, e2);
        u2 = ffma2(e2, d2, u2);
    }
#endif
}
// ---- exponentials on the FMA pipe (backward sweep of the fused pass) ----------------------------------------------
// Under the power cap the fused pass is bound by the SFU (one MUFU.EX2 per element and sweep, 16 per clock and SM),
// while the FMA pipe idles at 20 %.  B200TRL_K1_BWD_POLY8 of every 8 bf16 pairs of the BACKWARD sweep therefore take
// 2^d from a Cody-Waite split + cubic instead: n = round(d) by the 1.5 * 2^23 trick, r = d - n in [-0.5, 0.5],
// p(r) ~ 2^r (max rel. error 1.2e-4, a 30th of a bf16 half-ulp; the result is rounded to bf16 right after), and n is
// added into p's exponent field.  d <= 0 here (lse >= max); the logits are clamped from below (packed bf16 max, NaN
// propagating) so that d >= -125.5 and the exponent cannot wrap -- 2^-125 g is zero in bf16 anyway.  The forward sweep
// keeps MUFU: its sums need fp32-accurate terms.
#ifndef B200TRL_K1_BWD_POLY8
#define B200TRL_K1_BWD_POLY8 0
#endif
__device__ __forceinline__ uint32_t bf16x2_max_nan(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t d2) {
    const uint64_t magic2 = pack2(12582912.f, 12582912.f), nmagic2 = pack2(-12582912.f, -12582912.f);
    const uint64_t t2 = fadd2(d2, magic2);                            // low mantissa bits: round(d)
    const uint64_t r2 = fadd2(d2, fadd2(nmagic2, t2) ^ 0x8000000080000000ull);  // d - round(d)
    uint64_t p2 = ffma2(pack2(0.05507577f, 0.05507577f), r2, pack2(0.24237224f, 0.24237224f));
    p2 = ffma2(p2, r2, pack2(0.69325519f, 0.69325519f));
    p2 = ffma2(p2, r2, pack2(0.99994314f, 0.99994314f));
    float t0, t1, p0, p1;
    unpack2(t2, t0, t1);
    unpack2(p2, p0, p1);
    return pack2(__uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23)),
                 __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23)));
}
// one vector of dlogits with NP of its 4 pairs on the FMA pipe; xmin2: packed bf16 lower clamp of the row
template <int NP>
__device__ __forceinline__ uint4 grad_vec_mix(const uint4& v, uint64_t c2, uint64_t nl2, uint64_t ng2, uint32_t xmin2) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool poly = i < NP;
        const uint32_t wi = poly ? bf16x2_max_nan(w[i], xmin2) : w[i];
        const uint64_t x2 = pack2(__uint_as_float(wi << 16), __uint_as_float(wi & 0xffff0000u));
        const uint64_t d2 = ffma2(x2, c2, nl2);
        uint64_t e2;
        if (poly) {
            e2 = exp2_poly2(d2);
        } else {
            float d0, d1;
            unpack2(d2, d0, d1);
            e2 = pack2(ex2(d0), ex2(d1));
        }
        float o0, o1;
        unpack2(fmul2(e2, ng2), o0, o1);
        o[i] = cvt_bf16x2(o0, o1);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

template <int NC>
struct FusedConsumer {
    static constexpr int kChunkBytes = chunk_bytes_for(NC);
    static constexpr int kChunkElems = kChunkBytes / 2;
    static constexpr int kVpt = kChunkBytes / 16 / NC;
    static constexpr uint32_t kVecStride = NC * 16u;

    // geometry
    uint32_t slots_u32, full0, done0, my_off;
    int num_slots, tid, lane, tail_vecs;
    float c;
    uint64_t c2;
    // forward state of the row being folded
    int fslot;
    uint32_t fpar;
    float m;
    uint32_t rmx;
    uint64_t s2, u2, t2, v2;
    // backward cursor
    int bslot;

    __device__ __forceinline__ void row_begin() {
        m = kNegBig;
        rmx = kNegInf2;
        s2 = u2 = t2 = v2 = 0ull;
    }
    __device__ __forceinline__ Partial row_partial() const {
        float s0, s1, u0, u1;
        unpack2(fadd2(s2, t2), s0, s1);
        unpack2(fadd2(u2, v2), u0, u1);
        return Partial{m, s0 + s1, u0 + u1};
    }
    __device__ __forceinline__ void move_reference(uint32_t t) {  // the running maximum moved: rare after the first chunk
        rmx = t;
        const float cm = fmaxf(__uint_as_float(t << 16), __uint_as_float(t & 0xffff0000u)) * c;
        if (cm > m + kSlack) {
            const float d = m - cm;
            const float f = ex2(d);
            const uint64_t f2 = pack2(f, f), d2 = pack2(d, d);
            u2 = fmul2(f2, ffma2(d2, s2, u2));
            s2 = fmul2(s2, f2);
            v2 = fmul2(f2, ffma2(d2, t2, v2));
            t2 = fmul2(t2, f2);
            m = cm;
        }
    }
    __device__ __forceinline__ void fwd_advance() {
        if (++fslot == num_slots) {
            fslot = 0;
            fpar ^= 1u;
        }
    }
    // a chunk of a masked row (masked-row skipping): nothing was loaded, the DMA warp arrived on the barrier itself
    __device__ __forceinline__ void fwd_skip() {
        mbar_wait_u32(full0 + static_cast<uint32_t>(fslot) * 8u, fpar);
        fwd_advance();
    }
    // a full chunk: every thread folds kVpt vectors
    __device__ __forceinline__ void fwd_full() {
        mbar_wait_u32(full0 + static_cast<uint32_t>(fslot) * 8u, fpar);
        const uint32_t base = slots_u32 + static_cast<uint32_t>(fslot) * kChunkBytes + my_off;
        uint4 v[kVpt];
        uint32_t t = rmx;
#pragma unroll
        for (int k = 0; k < kVpt; ++k) {
            v[k] = lds128(base + k * kVecStride);
            t = bf16x2_max(t, vec_max(v[k]));
        }
        if (t != rmx) move_reference(t);
        const uint64_t nm2 = pack2(-m, -m);
#pragma unroll
        for (int k = 0; k < kVpt; k += 2) {
            fold_words(s2, u2, v[k], c2, nm2);
            fold_words(t2, v2, v[k + 1], c2, nm2);
        }
        fwd_advance();
    }
    // the partial last chunk of a slice: vector v of it belongs to thread v % NC; once per row
    __device__ __forceinline__ void fwd_tail() {
        mbar_wait_u32(full0 + static_cast<uint32_t>(fslot) * 8u, fpar);
        const uint32_t base = slots_u32 + static_cast<uint32_t>(fslot) * kChunkBytes + my_off;
        uint32_t t = rmx;
#pragma unroll
        for (int k = 0; k < kVpt; ++k)
            if (tid + k * NC < tail_vecs) t = bf16x2_max(t, vec_max(lds128(base + k * kVecStride)));
        if (t != rmx) move_reference(t);
        const uint64_t nm2 = pack2(-m, -m);
#pragma unroll
        for (int k = 0; k < kVpt; ++k) {
            if (tid + k * NC < tail_vecs) {
                const uint4 v = lds128(base + k * kVecStride);
                if (k & 1)
                    fold_words(t2, v2, v, c2, nm2);
                else
                    fold_words(s2, u2, v, c2, nm2);
            }
        }
        fwd_advance();
    }
    template <bool FULL>
    __device__ __forceinline__ void fwd() {
        if (FULL)
            fwd_full();
        else
            fwd_tail();
    }
    // one chunk of dlogits: gv = this CTA's slice of the row at this chunk
    template <bool FULL>
    __device__ __forceinline__ void bwd(uint4* gv, uint64_t nl2, uint64_t ng2, uint32_t xmin2) {
        constexpr int kPolyEven = (B200TRL_K1_BWD_POLY8 + 1) / 2, kPolyOdd = B200TRL_K1_BWD_POLY8 / 2;
        const uint32_t base = slots_u32 + static_cast<uint32_t>(bslot) * kChunkBytes + my_off;
        uint4 v[kVpt];
#pragma unroll
        for (int k = 0; k < kVpt; ++k) {
            v[k] = make_uint4(0u, 0u, 0u, 0u);
            if (FULL || (tid + k * NC < tail_vecs)) v[k] = lds128(base + k * kVecStride);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_u32(done0 + static_cast<uint32_t>(bslot) * 8u);  // read: the slot may be refilled
#pragma unroll
        for (int k = 0; k < kVpt; ++k)
            if (FULL || (tid + k * NC < tail_vecs))
                st_global_cs(gv + tid + k * NC, (k & 1) ? grad_vec_mix<kPolyOdd>(v[k], c2, nl2, ng2, xmin2)
                                                        : grad_vec_mix<kPolyEven>(v[k], c2, nl2, ng2, xmin2));
        if (++bslot == num_slots) bslot = 0;
    }
    template <bool FULL>
    __device__ __forceinline__ void bwd_zero(uint4* gv) {
        if (lane == 0) mbar_arrive_u32(done0 + static_cast<uint32_t>(bslot) * 8u);
#pragma unroll
        for (int k = 0; k < kVpt; ++k)
            if (FULL || (tid + k * NC < tail_vecs)) st_global_cs(gv + tid + k * NC, make_uint4(0u, 0u, 0u, 0u));
        if (++bslot == num_slots) bslot = 0;
    }
};

// In-kernel mask statistics (K1Args::count_mask): the consumer warps have nothing to do until the first chunk lands
// (~2 us), so they count the completion mask -- warp w the sequences w, w + n_warps, ... -- into shared memory; the
// reducer warp joins the named barrier before it reads them.  Replaces the mask_stats launch (+ its memset node) of a
// step.  Integer-valued fp32 sums: exact below 2^24 tokens, order-independent.
template <int NC>
__device__ __forceinline__ void count_mask_consumers(const K1Args& a, float* row_cnt, float* total_cnt, int warp, int lane) {
    constexpr int kWarps = NC / 32;
    const int B = static_cast<int>(a.B), T = static_cast<int>(a.T);
    for (int b = warp; b < B; b += kWarps) {
        const int32_t* m = a.mask + static_cast<int64_t>(b) * T;
        int n = 0;
        for (int t = lane; t < T; t += 32) n += m[t];
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0) row_cnt[b] = static_cast<float>(n);
    }
    asm volatile("bar.sync 2, %0;" ::"n"(NC) : "memory");  // consumers only: every row_cnt is written
    if (warp == 0) {
        float tot = 0.f;
        for (int b = lane; b < B; b += 32) tot += row_cnt[b];
        tot = warp_sum(tot);
        if (lane == 0) *total_cnt = tot;
    }
    asm volatile("bar.arrive 3, %0;" ::"n"(NC + 32) : "memory");  // hand over to the reducer warp (bar.sync 3)
}

// ------------------------------------------------------------------ the kernel
// Warp roles: 0..15 consumers, 16 DMA (one lane), 17 reducer (row statistics + cluster exchange).
//
// Consumer schedule for row i (fused mode), software-pipelined so that the reduce / DSMEM exchange latency of
// row i hides behind useful work:
//     forward chunks [k_pre, C) of row i          (chunks [0, k_pre) were folded one step earlier)
//     publish the warp partial of row i            -> reducer
//     forward chunks [0, k_pre) of row i+1         (already prefetched into the spare slots)
//     wait for RowResult(i)                        <- reducer
//     backward chunks [0, C) of row i              -> DMA stores, slots refilled with row i+1 / i+2
struct Cursor {
    int slot;
    uint32_t par;
    __device__ __forceinline__ void advance(int num_slots) {
        if (++slot == num_slots) {
            slot = 0;
            par ^= 1u;
        }
    }
};

// FAST: 0 = generic consumer code; 1 = FusedConsumer
template <bool HAS_FWD, bool HAS_BWD, bool DUAL, int NC, bool DIRECT, bool SKIP, int FAST, bool SKEW, bool F16>
__global__ void __launch_bounds__(NC + 64, (NC <= 256) ? 2 : 1)  // 768 consumers: <= 78 registers
    k1_resident_kernel(const K1Args a, const int num_slots, const int max_lag, const int l2_prefetch) {
    static_assert(!F16 || FAST == 0, "fp16 logits run on the generic consumer code");
    using ElemT = typename std::conditional<F16, __half, __nv_bfloat16>::type;
    static_assert(FAST == 0 || (HAS_FWD && HAS_BWD && DUAL && DIRECT), "the fast consumer is the fused pass");
    static_assert(!SKEW || (FAST == 0 && (DIRECT || !HAS_BWD)), "skewed rows: generic consumers, dlogits from registers");
    constexpr int kConsumers = NC;
    constexpr int kWarps = NC / 32;
    // chunk geometry of this instantiation: 512 and 256 consumers use 16 KB chunks (2 / 4 vectors per thread),
    // 768 consumers use 24 KB chunks (2 vectors per thread)
    constexpr int kChunkBytes = chunk_bytes_for(NC);
    constexpr int kChunkElems = kChunkBytes / 2;
    constexpr int kChunkVecs = kChunkBytes / 16;
    constexpr int kVpt = kChunkVecs / NC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* slots = smem_raw;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw + static_cast<size_t>(num_slots) * kChunkBytes);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_rank();
    const uint32_t csize = cluster_size();
    const int64_t first_row = cluster_id_x();
    const int64_t row_step = num_clusters_x();

    // this CTA's slice of every row
    const int64_t slice_elems = ((a.vocab + csize - 1) / csize + 7) & ~int64_t(7);
    const int64_t e_begin = static_cast<int64_t>(crank) * slice_elems;
    const int64_t e_end = min(a.vocab, e_begin + slice_elems);
    const int64_t my_elems = max(e_end - e_begin, (int64_t)0);
    // SKEW: the aligned span of a slice is up to 7 elements longer at either end; C covers the longest one and a
    // row's last chunk may turn out empty (the DMA lane then completes its barrier without a copy)
    const int my_bytes = SKEW ? static_cast<int>(((my_elems + 14) >> 3) << 4) : static_cast<int>(my_elems * 2);
    const int C = (my_bytes + kChunkBytes - 1) / kChunkBytes;  // chunks per row in this CTA
    // head skew of a row's slice (elements between the 16-byte boundary below it and its first element) and the
    // number of 16-byte vectors of its aligned span
    const uint64_t base_elem = static_cast<uint64_t>(reinterpret_cast<uintptr_t>(a.logits) >> 1);
    auto row_head = [&](int64_t row) -> int {
        return SKEW ? static_cast<int>((base_elem + static_cast<uint64_t>(logits_offset(a, row) + e_begin)) & 7u) : 0;
    };
    auto span_vecs = [&](int h) -> int { return SKEW ? static_cast<int>((h + my_elems + 7) >> 3) : (my_bytes >> 4); };
    const int n_my_rows =
        (a.n_rows > first_row) ? static_cast<int>((a.n_rows - first_row + row_step - 1) / row_step) : 0;
    const int spare = max(num_slots - C, 0);
    const int k_pre = (HAS_FWD && HAS_BWD) ? min(spare, C) : 0;  // chunks of the next row folded early
    const int last_bytes = my_bytes - (C - 1) * kChunkBytes;      // bytes of the last chunk of a slice

    if (tid == 0) {
        for (int s = 0; s < num_slots; ++s) {
            mbar_init(&sm.full_bar[s], 1);
            mbar_init(&sm.done_bar[s], kWarps);
        }
        for (int p = 0; p < 2; ++p) {
            mbar_init(&sm.part_bar[p], kWarps);
            mbar_init(&sm.res_bar[p], 1);
            mbar_init(&sm.xchg_bar[p], 1);  // one local arrive.expect_tx per row; peers complete the bytes
        }
        fence_barrier_init();
    }
    if (a.gmode == G_PPO && warp == kWarps) {
        const float n = ppo_unpadded_count(a, lane);
        const float n1 = a.step_ws ? ppo_unpadded_count(a, lane, 1) : 0.f;
        if (lane == 0) {
            sm.ppo_count = n;
            sm.ppo_count_p1 = n1;
        }
    }
    __syncthreads();
    if (csize > 1) cluster_sync_all();  // peers' barriers are initialised before anyone signals them

    // 16-bit elements either way: the pointer arithmetic (element offsets) is the same for bf16 and fp16
    const __nv_bfloat16* logits = reinterpret_cast<const __nv_bfloat16*>(a.logits);
    __nv_bfloat16* dlogits = reinterpret_cast<__nv_bfloat16*>(a.dlogits);

    if (warp == kWarps) {
        // =========================== DMA warp: one lane drives every bulk copy ===========================
        if (lane == 0 && n_my_rows > 0 && C > 0) {
            const uint64_t policy = policy_evict_first();
            const int64_t J = static_cast<int64_t>(n_my_rows) * C;
            Tracer tr;
            tr.init(a, &sm.trace, 2, true);
            // load cursor
            int64_t k_next = 0;
            int l_row = 0, l_c = 0, l_slot = 0;
            // masked-row skipping: the flag of the row being loaded and (prefetched) of the one after it
            auto masked_at = [&](int r) {
                return SKIP && r < n_my_rows && row_is_masked(a, first_row + static_cast<int64_t>(r) * row_step);
            };
            bool l_masked = masked_at(0), l_masked_next = masked_at(1);
            auto issue_load = [&]() {
                const int64_t row = first_row + static_cast<int64_t>(l_row) * row_step;
                uint32_t bytes = static_cast<uint32_t>(l_c == C - 1 ? last_bytes : kChunkBytes);
                int l_h = 0;
                if (SKEW) {  // this row's span: the chunk may be short or empty
                    l_h = row_head(row);
                    const int rest = span_vecs(l_h) * 16 - l_c * kChunkBytes;
                    bytes = static_cast<uint32_t>(min(max(rest, 0), kChunkBytes));
                }
                if (l2_prefetch && l_c == 0 && l_row + 1 < n_my_rows && !(SKIP && l_masked_next)) {
                    // pull the NEXT row's slice into L2 now: its shared-memory slots only free up while this row is
                    // being written back, and a bulk load that hits L2 lands in a fraction of the HBM queueing time
                    const __nv_bfloat16* nxt = logits + logits_offset(a, row + row_step) + e_begin;
                    for (int cc = 0; cc < C; ++cc)
                        bulk_prefetch_l2(nxt + static_cast<int64_t>(cc) * kChunkElems,
                                         static_cast<uint32_t>(cc == C - 1 ? last_bytes : kChunkBytes));
                }
                if ((SKIP && l_masked) || (SKEW && bytes == 0)) {
                    mbar_arrive(&sm.full_bar[l_slot]);  // nothing to fetch: the slot is "full" right away
                } else {
                    tr.ev(21, l_row, l_c);
                    mbar_expect_tx(&sm.full_bar[l_slot], bytes);
                    bulk_load(slots + static_cast<size_t>(l_slot) * kChunkBytes,
                              logits + logits_offset(a, row) + e_begin - l_h + static_cast<int64_t>(l_c) * kChunkElems, bytes,
                              &sm.full_bar[l_slot], policy);
                }
                ++k_next;
                if (++l_c == C) {
                    l_c = 0;
                    ++l_row;
                    l_masked = l_masked_next;
                    l_masked_next = masked_at(l_row + 1);
                }
                if (++l_slot == num_slots) l_slot = 0;
            };
            while (k_next < J && k_next < num_slots) issue_load();
            constexpr bool kBulkStore = HAS_BWD && !DIRECT;  // dlogits leave through TMA bulk stores (else: consumers' STG)
            const int lag = kBulkStore ? min(max_lag, spare) : 0;  // stores allowed to be still reading their slot
            Cursor cur{0, 0u};
            int s_row = 0, s_c = 0;
            int64_t drained = 0;  // slots [0, drained) of the chunk stream are free again
            for (int64_t j = 0; j < J; ++j) {
                mbar_wait(&sm.done_bar[cur.slot], cur.par);
                tr.ev(20, static_cast<int>(j / C), static_cast<int>(j % C));
                if (kBulkStore) {
                    const int64_t row = first_row + static_cast<int64_t>(s_row) * row_step;
                    const uint32_t bytes = static_cast<uint32_t>(s_c == C - 1 ? last_bytes : kChunkBytes);
                    bulk_store(dlogits + dlogits_offset(a, row) + e_begin + static_cast<int64_t>(s_c) * kChunkElems,
                               slots + static_cast<size_t>(cur.slot) * kChunkBytes, bytes);
                    bulk_commit();
                    const bool row_end = (s_c == C - 1);
                    if (row_end || lag == 0) {
                        bulk_wait_read<0>();
                        drained = j + 1;
                    } else if (lag == 1) {
                        bulk_wait_read<1>();
                        drained = j;
                    } else {
                        bulk_wait_read<2>();
                        drained = j - 1;
                    }
                    tr.ev(22, s_row, s_c);  // store issued and the allowed lag drained
                    if (++s_c == C) {
                        s_c = 0;
                        ++s_row;
                    }
                } else {
                    drained = j + 1;
                }
                while (k_next < J && k_next - num_slots < drained) issue_load();
                cur.advance(num_slots);
            }
            if (kBulkStore) bulk_wait_all();
            tr.finish();
        }
    } else if (warp == kWarps + 1) {
        // =========================== reducer warp: row statistics, cluster exchange, token gradient ===========
        if (HAS_FWD) {
            const float c = a.c;
            uint32_t xphase[2] = {0u, 0u};  // phase parity of the two exchange barriers (one flip per use)
            // Row scalars are fetched two rows ahead (lane 0): the id-addressed logit of row i+1 and the directly
            // addressed scalars of row i+2 are issued before row i is processed, so neither of the two dependent
            // global-load latencies sits on a row's hand-off.
            const float ppo_count = (a.gmode == G_PPO) ? sm.ppo_count : 1.f;
            Tracer tr;
            tr.init(a, &sm.trace, 1, lane == 0);
            // in-kernel loss / metric sums of this cluster's rows (leader CTA, lane 0): loss, kl, entropy, low, high, region
            const bool step_sums = HAS_BWD && a.step_ws != nullptr && a.gmode == G_GRPO && crank == 0;
            float sums[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            // same for the PPO micro-batch statistics (+ d loss / d vpred, one float per row)
            const bool ppo_sums = HAS_BWD && a.step_ws != nullptr && a.gmode == G_PPO && crank == 0;
            float psums[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            auto ppo_row = [&](int64_t row, const RowScalars& r, float logp, float entropy) {  // lane 0, after the publish
                const int64_t b = row / a.T, t = row - b * a.T;
                const bool pad1 = t > a.seq_len[b] + 1;
                const float dv = ppo_token_stats(logp, r.aux0, r.adv, a.ppo_vpred[row], a.ppo_values[row],
                                                 a.ppo_returns[row], entropy, r.pad != 0.f, pad1, a.clip_lo, a.clip_hi,
                                                 a.cliprange_value, psums);
                if (a.ppo_dvpred) a.ppo_dvpred[row] = dv * (0.5f * a.vf_coef * a.grad_scale / sm.ppo_count_p1);
            };
            const bool counted = HAS_BWD && a.count_mask != 0;
            if (counted) asm volatile("bar.sync 3, %0;" ::"n"(NC + 32) : "memory");  // the consumers counted the mask
            const float* row_count = counted ? sm.row_cnt : a.row_count;
            const float* total_count = counted ? &sm.total_cnt : a.total_count;
            RowScalars rs{}, rs_next{};
            if (lane == 0 && n_my_rows > 0) {
                rs = load_row_scalars_direct(a, first_row, ppo_count, row_count, total_count);
                rs.x_sel = load_selected_logit<ElemT>(a, first_row, rs.id);
                if (n_my_rows > 1) rs_next = load_row_scalars_direct(a, first_row + row_step, ppo_count, row_count, total_count);
            }
            for (int i = 0; i < n_my_rows; ++i) {
                const int64_t row = first_row + static_cast<int64_t>(i) * row_step;
                const int par = i & 1;
                const uint32_t rpar = static_cast<uint32_t>((i >> 1) & 1);
                const bool skip_row = SKIP && row_is_masked(a, row);
                float x_next = 0.f;
                RowScalars rs_after{};
                if (lane == 0) {
                    if (i + 1 < n_my_rows) x_next = load_selected_logit<ElemT>(a, row + row_step, rs_next.id);
                    if (i + 2 < n_my_rows)
                        rs_after = load_row_scalars_direct(a, row + 2 * row_step, ppo_count, row_count, total_count);
                }
                tr.ev(10, i, 0);
                mbar_wait(&sm.part_bar[par], rpar);
                tr.ev(11, i, 0);
                if (skip_row) {  // nothing was read: outputs are zero, the row's dlogits are zero
                    if (lane == 0) {
                        if (crank == 0) {
                            if (a.logp) a.logp[row] = (a.gmode == G_PPO) ? 1.0f : 0.f;
                            if (a.entropy) a.entropy[row] = 0.f;
                            if (a.lse) a.lse[row] = 0.f;
                        }
                        RowResult rr;
                        rr.lse2 = 0.f;
                        rr.ng = 0.f;
                        rr.patch = 0.f;
                        rr.id_chunk = rr.id_vec = rr.id_elem = -1;
                        rr.pad0 = rr.pad1 = 0;
                        sm.result[par] = rr;
                        mbar_arrive(&sm.res_bar[par]);
                        if (ppo_sums) ppo_row(row, rs, 1.0f, 0.f);  // a skipped row is a pad row: INVALID_LOGPROB, entropy 0
                    }
                    __syncwarp();
                    rs = rs_next;
                    rs.x_sel = x_next;
                    rs_next = rs_after;
                    continue;
                }
                Partial q = partial_empty();
                if (lane < kWarps) {
                    const Part4 w = sm.warp_part[par][lane];
                    q = Partial{w.m, w.s, w.u};
                }
                q = partial_warp_reduce_fast(q);  // every lane holds the CTA's partial
                if (csize > 1) {
                    // lane r posts this CTA's partial to CTA r (itself included); the bytes complete the receiver's
                    // barrier, so the only latency on the row's hand-off is one DSMEM write
                    if (lane == 0) mbar_expect_tx(&sm.xchg_bar[par], 16u * csize);
                    if (lane < csize)
                        st_async_v4(map_to_rank(smem_u32(&sm.xchg[par][crank]), lane), q.m, q.s, q.u, 0.f,
                                    map_to_rank(smem_u32(&sm.xchg_bar[par]), lane));
                }
                if (lane == 0) {
                    if (csize > 1) {
                        tr.ev(12, i, 0);
                        mbar_wait(&sm.xchg_bar[par], xphase[par]);
                        tr.ev(13, i, 0);
                        xphase[par] ^= 1u;
                        Partial tot = partial_empty();
                        for (uint32_t r = 0; r < csize; ++r) {  // rank order: identical result in every CTA
                            const Part4 w = sm.xchg[par][r];
                            tot = partial_merge(tot, Partial{w.m, w.s, w.u});
                        }
                        q = tot;
                    }
                    const RowStats st = finish_row(q, rs.x_sel, c);
                    if (HAS_BWD) {
                        GrpoTok tok;
                        tok.loss = tok.kl = tok.low = tok.high = 0.f;
                        const float gp = token_grad(a, rs, st.logp, &tok) * a.inv_temp;
                        if (step_sums && rs.norm != 0.f) {  // a token the loss sees (mask == 1), in this CTA's row order
                            sums[0] = fmaf(tok.loss, rs.norm, sums[0]);
                            sums[1] += tok.kl;
                            sums[2] += st.entropy;
                            sums[3] += tok.low;
                            sums[4] += tok.high;
                            sums[5] += fmaxf(tok.low, tok.high);
                        }
                        const int64_t e_id = rs.id - e_begin;
                        const bool mine = (e_id >= 0 && e_id < my_elems);
                        const int64_t s_id = e_id + row_head(row);  // position inside the aligned span
                        RowResult rr;
                        rr.lse2 = st.lse2;
                        rr.ng = -gp;
                        rr.patch = fmaf(-expf(st.logp), gp, gp);  // g' * (1 - p_id)
                        rr.id_chunk = mine ? static_cast<int>(s_id / kChunkElems) : -1;
                        rr.id_elem = mine ? static_cast<int>(s_id % kChunkElems) : -1;
                        rr.id_vec = mine ? (rr.id_elem >> 3) : -1;
                        rr.pad0 = rr.pad1 = 0;
                        sm.result[par] = rr;
                    }
                    mbar_arrive(&sm.res_bar[par]);  // the consumers are waiting for this: publish first, store after
                    tr.ev(14, i, 0);
                    if (crank == 0) {
                        const bool pad = (a.gmode == G_PPO) && rs.pad != 0.f;
                        if (a.logp) a.logp[row] = pad ? 1.0f : st.logp;
                        if (a.entropy) a.entropy[row] = st.entropy;
                        if (a.lse) a.lse[row] = st.lse;
                        if (ppo_sums) ppo_row(row, rs, st.logp, st.entropy);
                    }
                }
                __syncwarp();
                rs = rs_next;
                rs.x_sel = x_next;
                rs_next = rs_after;
            }
            if (HAS_BWD && a.step_ws != nullptr && a.gmode == G_PPO && crank == 0) {
                unsigned int* counter = reinterpret_cast<unsigned int*>(a.step_ws);
                volatile float* part = a.step_ws + 4;
                const unsigned int n_cl = num_clusters_x();
                bool last = false;
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 7; ++k) part[static_cast<size_t>(cluster_id_x()) * 8 + k] = psums[k];
                    __threadfence();
                    const unsigned int prev = atomicAdd(counter, 1u);
                    last = (prev == n_cl - 1);
                    if (last) {
                        *counter = 0u;
                        __threadfence();
                    }
                }
                last = __shfl_sync(0xffffffffu, last ? 1 : 0, 0) != 0;
                if (last) {
                    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
                    for (unsigned int cl = lane; cl < n_cl; cl += 32)
#pragma unroll
                        for (int k = 0; k < 7; ++k) acc[k] += static_cast<double>(part[static_cast<size_t>(cl) * 8 + k]);
#pragma unroll
                    for (int k = 0; k < 7; ++k) acc[k] = warp_sum(acc[k]);
                    if (lane == 0) {  // ppo_trainer.py:573-605, as b200trl_ppo_loss writes them
                        const double n_pad = sm.ppo_count, n_p1 = sm.ppo_count_p1;
                        const double n_all = static_cast<double>(a.B) * static_cast<double>(a.T);
                        const double pg_loss = acc[0] / n_pad, vf_loss = 0.5 * acc[1] / n_p1;
                        a.step_metrics[B200TRL_P_LOSS] = static_cast<float>(pg_loss + a.vf_coef * vf_loss);
                        a.step_metrics[B200TRL_P_PG_LOSS] = static_cast<float>(pg_loss);
                        a.step_metrics[B200TRL_P_VF_LOSS] = static_cast<float>(vf_loss);
                        a.step_metrics[B200TRL_P_PG_CLIPFRAC] = static_cast<float>(acc[2] / n_pad);
                        a.step_metrics[B200TRL_P_VF_CLIPFRAC] = static_cast<float>(acc[3] / n_p1);
                        a.step_metrics[B200TRL_P_APPROXKL] = static_cast<float>(0.5 * acc[4] / n_all);
                        a.step_metrics[B200TRL_P_ENTROPY] = static_cast<float>(acc[5] / n_all);
                        a.step_metrics[B200TRL_P_RATIO] = static_cast<float>(acc[6] / n_all);
                    }
                }
            }
            if (HAS_BWD && a.step_ws != nullptr && a.gmode == G_GRPO && crank == 0) {
                // leave this cluster's sums; the last cluster to arrive folds all of them in cluster order (double):
                // deterministic for a given grid, one launch
                unsigned int* counter = reinterpret_cast<unsigned int*>(a.step_ws);
                volatile float* part = a.step_ws + 4;
                const unsigned int n_cl = num_clusters_x();
                bool last = false;
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) part[static_cast<size_t>(cluster_id_x()) * 8 + k] = sums[k];
                    __threadfence();
                    const unsigned int prev = atomicAdd(counter, 1u);
                    last = (prev == n_cl - 1);
                    if (last) {
                        *counter = 0u;  // ready for the next launch
                        __threadfence();
                    }
                }
                last = __shfl_sync(0xffffffffu, last ? 1 : 0, 0) != 0;
                if (last) {
                    double acc[6] = {0, 0, 0, 0, 0, 0};
                    for (unsigned int cl = lane; cl < n_cl; cl += 32)
#pragma unroll
                        for (int k = 0; k < 6; ++k) acc[k] += static_cast<double>(part[static_cast<size_t>(cl) * 8 + k]);
#pragma unroll
                    for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
                    if (lane == 0) {
                        const float total = total_count[0];
                        const double ntok = fmax(static_cast<double>(total), 1.0);
                        a.step_loss[0] = static_cast<float>(acc[0]);
                        a.step_metrics[B200TRL_M_LOSS] = static_cast<float>(acc[0]);
                        a.step_metrics[B200TRL_M_KL] = static_cast<float>(acc[1] / ntok);
                        a.step_metrics[B200TRL_M_ENTROPY] = static_cast<float>(acc[2] / ntok);
                        a.step_metrics[B200TRL_M_CLIP_LOW] = static_cast<float>(acc[3] / ntok);
                        a.step_metrics[B200TRL_M_CLIP_HIGH] = static_cast<float>(acc[4] / ntok);
                        a.step_metrics[B200TRL_M_CLIP_REGION] = static_cast<float>(acc[5] / ntok);
                        a.step_metrics[B200TRL_M_TOKENS] = total;
                        a.step_metrics[B200TRL_M_RESERVED] = 0.f;
                    }
                }
            }
            tr.finish();
        }
    } else if (FAST != 0) {
        // =========================== consumers, fused pass (FusedConsumer) ===========================
        using FC = FusedConsumer<NC>;
        if (a.count_mask != 0) count_mask_consumers<NC>(a, sm.row_cnt, &sm.total_cnt, warp, lane);
        FC fc;
        fc.slots_u32 = smem_u32(slots);
        fc.full0 = smem_u32(&sm.full_bar[0]);
        fc.done0 = smem_u32(&sm.done_bar[0]);
        fc.my_off = static_cast<uint32_t>(tid) * 16u;
        fc.num_slots = num_slots;
        fc.tid = tid;
        fc.lane = lane;
        fc.c = a.c;
        fc.c2 = pack2(a.c, a.c);
        fc.fslot = 0;
        fc.fpar = 0u;
        fc.bslot = 0;
        const int n_full = (last_bytes == kChunkBytes) ? C : C - 1;  // chunks every thread has kVpt vectors of
        fc.tail_vecs = (n_full < C) ? (last_bytes >> 4) : 0;
        const int pre_full = min(k_pre, n_full);
        const bool pre_tail = (k_pre == C) && (n_full < C);    // the early-folded chunks include the partial one
        const bool main_tail = (k_pre < C) && (n_full < C);
        // masked-row skipping (SKIP): the flag of the row being folded and of the one after it, fetched ahead
        auto masked_at = [&](int r) {
            return SKIP && r < n_my_rows && row_is_masked(a, first_row + static_cast<int64_t>(r) * row_step);
        };
        bool cur_masked = masked_at(0), nxt_masked = masked_at(1);
        auto fold_early = [&](bool masked) {  // chunks [0, k_pre)
            if (SKIP && masked) {
                for (int cidx = 0; cidx < k_pre; ++cidx) fc.fwd_skip();
            } else {
                for (int cidx = 0; cidx < pre_full; ++cidx) fc.template fwd<true>();
                if (pre_tail) fc.template fwd<false>();
            }
        };
        fc.row_begin();
        if (n_my_rows > 0) fold_early(cur_masked);
        for (int i = 0; i < n_my_rows; ++i) {
            const int64_t row = first_row + static_cast<int64_t>(i) * row_step;
            const int par = i & 1;
            const uint32_t rpar = static_cast<uint32_t>((i >> 1) & 1);
            const bool nn_masked = masked_at(i + 2);
            if (SKIP && cur_masked) {
                for (int cidx = k_pre; cidx < C; ++cidx) fc.fwd_skip();
            } else {
                for (int cidx = k_pre; cidx < n_full; ++cidx) fc.template fwd<true>();
                if (main_tail) fc.template fwd<false>();
            }
            const Partial p = partial_warp_reduce_fast(fc.row_partial());
            if (lane == 0) {
                sm.warp_part[par][warp] = Part4{p.m, p.s, p.u, 0.f};
                mbar_arrive(&sm.part_bar[par]);
            }
            fc.row_begin();
            if (i + 1 < n_my_rows) fold_early(nxt_masked);
            cur_masked = nxt_masked;
            nxt_masked = nn_masked;
            mbar_wait(&sm.res_bar[par], rpar);
            const RowResult rr = sm.result[par];
            const uint64_t nl2 = pack2(-rr.lse2, -rr.lse2);
            const uint64_t ng2 = pack2(rr.ng, rr.ng);
            uint4* gv = reinterpret_cast<uint4*>(dlogits + dlogits_offset(a, row) + e_begin);
            if (rr.ng == 0.f) {  // masked token: the row's dlogits are zero
                for (int cidx = 0; cidx < n_full; ++cidx) fc.template bwd_zero<true>(gv + static_cast<int64_t>(cidx) * (kChunkBytes / 16));
                if (n_full < C) fc.template bwd_zero<false>(gv + static_cast<int64_t>(n_full) * (kChunkBytes / 16));
            } else {
                // lower clamp of the logits for the FMA-pipe exponentials: x c - lse2 >= -125.5 (bf16 rounding included)
                const uint32_t xm = __float_as_uint((rr.lse2 - 125.f) / fc.c) >> 16;
                const uint32_t xmin2 = xm | (xm << 16);
                for (int cidx = 0; cidx < n_full; ++cidx)
                    fc.template bwd<true>(gv + static_cast<int64_t>(cidx) * (kChunkBytes / 16), nl2, ng2, xmin2);
                if (n_full < C)
                    fc.template bwd<false>(gv + static_cast<int64_t>(n_full) * (kChunkBytes / 16), nl2, ng2, xmin2);
                // the selected id: g' * (1 - p_id); the thread that wrote the vector holding it patches it
                if (rr.id_vec >= 0 && (rr.id_vec % kConsumers) == tid)
                    reinterpret_cast<__nv_bfloat16*>(gv + static_cast<int64_t>(rr.id_chunk) * (kChunkBytes / 16))[rr.id_elem] =
                        __float2bfloat16_rn(rr.patch);
            }
        }
    } else {
        // =========================== consumers ===========================
        if (HAS_FWD && HAS_BWD && a.count_mask != 0) count_mask_consumers<NC>(a, sm.row_cnt, &sm.total_cnt, warp, lane);
        const float c = a.c;
        const uint64_t c2 = pack2(c, c);
        Cursor fcur{0, 0u}, bcur{0, 0u};
        Tracer tr;
        tr.init(a, &sm.trace, 0, tid == 0);
        int t_row = 0;  // row whose chunk the next fwd_chunk call folds (tracing only)

        // SKEW: head skew / span of the row being folded (f_*) and of the row whose dlogits are being written (b_*)
        // per row (hoisted out of the chunk loops): chunk holding the span's last vector and the valid elements of that
        // vector (0: all eight)
        int f_h = 0, f_nvec = 0, f_last = 0, f_keep = 0;
        const int tail_keep_of_h0 = static_cast<int>(my_elems & 7);  // (h + my_elems) & 7 with h added per row
        auto set_fwd_row = [&](int64_t r) {
            f_h = row_head(r);
            f_nvec = span_vecs(f_h);
            f_last = (f_nvec - 1) / kChunkVecs;
            f_keep = (f_h + tail_keep_of_h0) & 7;
        };
        const bool want_ent = a.entropy != nullptr;  // forward-only: the full chunks skip the entropy sum when nobody reads it
        auto fwd_chunk = [&](Acc& acc, int cidx, bool skip) {
            tr.ev(1, t_row, cidx);
            mbar_wait(&sm.full_bar[fcur.slot], fcur.par);
            tr.ev(2, t_row, cidx);
            const uint4* sv = reinterpret_cast<const uint4*>(slots + static_cast<size_t>(fcur.slot) * kChunkBytes);
            if (SKIP && skip) {
                // masked row: nothing was loaded into this slot
            } else if (SKEW && cidx < f_last && (cidx != 0 || f_h == 0)) {  // an interior chunk: nothing to mask
                uint4 v[kVpt];
#pragma unroll
                for (int k = 0; k < kVpt; ++k) v[k] = sv[tid + k * kConsumers];
                if (!HAS_BWD && !want_ent) {
#pragma unroll
                    for (int k = 0; k < kVpt; k += 2) acc_vec2<F16, false>(acc, v[k], v[k + 1], c, c2);
                } else {
#pragma unroll
                    for (int k = 0; k < kVpt; k += 2) acc_vec2<F16>(acc, v[k], v[k + 1], c, c2);
                }
            } else if (SKEW) {
                const int v0 = cidx * kChunkVecs;
                const int n_here = min(max(f_nvec - v0, 0), kChunkVecs);
                const int keep = f_keep;
                const bool edge = (cidx == 0 && f_h != 0) || (keep != 0 && cidx == f_last);
                if (n_here == kChunkVecs) {
                    uint4 v[kVpt];
#pragma unroll
                    for (int k = 0; k < kVpt; ++k) v[k] = sv[tid + k * kConsumers];
                    if (edge) {  // a full chunk holding the span's first or last vector: only that vector is masked
#pragma unroll
                        for (int k = 0; k < kVpt; ++k) {
                            const int gi = v0 + tid + k * kConsumers;
                            const int lo = (gi == 0) ? f_h : 0;
                            const int hi = (gi == f_nvec - 1 && keep != 0) ? keep : 8;
                            if (lo != 0 || hi != 8) v[k] = mask_vec<F16>(v[k], lo, hi);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kVpt; k += 2) acc_vec2<F16>(acc, v[k], v[k + 1], c, c2);
                } else {
                    for (int v = tid; v < n_here; v += kConsumers) {
                        uint4 x = sv[v];
                        const int lo = (v0 + v == 0) ? f_h : 0;
                        const int hi = (v0 + v == f_nvec - 1 && keep != 0) ? keep : 8;
                        if (lo != 0 || hi != 8) x = mask_vec<F16>(x, lo, hi);
                        acc_vec<F16>(acc, x, c, c2);
                    }
                }
            } else if (cidx != C - 1 || last_bytes == kChunkBytes) {
                uint4 v[kVpt];
#pragma unroll
                for (int k = 0; k < kVpt; ++k) v[k] = sv[tid + k * kConsumers];
#pragma unroll
                for (int k = 0; k < kVpt; k += 2) {
                    if (DUAL && !HAS_BWD && !want_ent) {  // forward-only pass without entropies: no U accumulation
                        acc_vec2<F16, false>(acc, v[k], v[k + 1], c, c2);
                    } else if (DUAL) {
                        acc_vec2<F16>(acc, v[k], v[k + 1], c, c2);
                    } else {
                        acc_vec<F16>(acc, v[k], c, c2);
                        acc_vec<F16>(acc, v[k + 1], c, c2);
                    }
                }
            } else {
                const int nvec = last_bytes >> 4;
                for (int v = tid; v < nvec; v += kConsumers) acc_vec<F16>(acc, sv[v], c, c2);
            }
            if (!HAS_BWD) {  // forward only: the slot can be refilled as soon as every warp has read it
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.done_bar[fcur.slot]);
            }
            tr.ev(3, t_row, cidx);
            fcur.advance(num_slots);
        };

        Acc acc = acc_empty();
        auto masked_at = [&](int r) {
            return SKIP && r < n_my_rows && row_is_masked(a, first_row + static_cast<int64_t>(r) * row_step);
        };
        bool cur_masked = masked_at(0), nxt_masked = masked_at(1);
        if (SKEW && n_my_rows > 0) set_fwd_row(first_row);
        if (HAS_FWD && n_my_rows > 0) {
            for (int cidx = 0; cidx < k_pre; ++cidx) fwd_chunk(acc, cidx, cur_masked);
        }
        for (int i = 0; i < n_my_rows; ++i) {
            const int64_t row = first_row + static_cast<int64_t>(i) * row_step;
            const int par = i & 1;
            const uint32_t rpar = static_cast<uint32_t>((i >> 1) & 1);
            const bool nn_masked = masked_at(i + 2);  // prefetched: consumed two iterations later
            const int b_h = SKEW ? row_head(row) : 0;
            const int b_nvec = span_vecs(b_h);

            if (HAS_FWD) {
                t_row = i;
                if (SKEW) set_fwd_row(row);
                for (int cidx = k_pre; cidx < C; ++cidx) fwd_chunk(acc, cidx, cur_masked);
                const Partial p = partial_warp_reduce_fast(acc_to_partial(acc));
                tr.ev(4, i, 0);
                if (!HAS_BWD && i >= 2) mbar_wait(&sm.res_bar[par], static_cast<uint32_t>(((i - 2) >> 1) & 1));
                if (lane == 0) {
                    sm.warp_part[par][warp] = Part4{p.m, p.s, p.u, 0.f};
                    mbar_arrive(&sm.part_bar[par]);
                }
                tr.ev(5, i, 0);
                acc = acc_empty();
                t_row = i + 1;
                if (i + 1 < n_my_rows) {
                    if (SKEW) set_fwd_row(row + row_step);
                    for (int cidx = 0; cidx < k_pre; ++cidx) fwd_chunk(acc, cidx, nxt_masked);
                }
            }
            cur_masked = nxt_masked;
            nxt_masked = nn_masked;

            if (HAS_BWD) {
                RowResult rr;
                if (HAS_FWD) {
                    tr.ev(6, i, 0);
                    mbar_wait(&sm.res_bar[par], rpar);
                    tr.ev(7, i, 0);
                    rr = sm.result[par];
                } else {
                    if (tid == 0) {
                        const RowScalars rs = load_row_scalars<ElemT>(a, row, a.gmode == G_PPO ? sm.ppo_count : 1.f);
                        const float lse2 = a.lse_in[row] * kLog2e;
                        const float logp = fmaf(rs.x_sel, c, -lse2) * kLn2;
                        const float gp = token_grad(a, rs, logp) * a.inv_temp;
                        const int64_t e_id = rs.id - e_begin;
                        const bool mine = (e_id >= 0 && e_id < my_elems);
                        const int64_t s_id = e_id + row_head(row);
                        RowResult w;
                        w.lse2 = lse2;
                        w.ng = -gp;
                        w.patch = fmaf(-expf(logp), gp, gp);
                        w.id_chunk = mine ? static_cast<int>(s_id / kChunkElems) : -1;
                        w.id_elem = mine ? static_cast<int>(s_id % kChunkElems) : -1;
                        w.id_vec = mine ? (w.id_elem >> 3) : -1;
                        w.pad0 = w.pad1 = 0;
                        sm.result[par] = w;
                    }
                    consumer_bar<NC>();
                    rr = sm.result[par];
                }
                const uint64_t nl2 = pack2(-rr.lse2, -rr.lse2);
                const uint64_t ng2 = pack2(rr.ng, rr.ng);
                const bool zero_row = (rr.ng == 0.f);
                const bool patch_mine = (rr.id_vec >= 0) && ((rr.id_vec % kConsumers) == tid);
                __nv_bfloat16* grow = nullptr;  // DIRECT: this CTA's slice of the dlogits row (SKEW: its aligned span)
                if (DIRECT) grow = dlogits + dlogits_offset(a, row) + e_begin - b_h;
                const int b_keep = SKEW ? ((b_h + tail_keep_of_h0) & 7) : 0;
                const int b_last = SKEW ? (b_nvec - 1) / kChunkVecs : 0;
                for (int cidx = 0; cidx < C; ++cidx) {
                    if (!HAS_FWD) mbar_wait(&sm.full_bar[bcur.slot], bcur.par);
                    uint4* sv = reinterpret_cast<uint4*>(slots + static_cast<size_t>(bcur.slot) * kChunkBytes);
                    bool full = (cidx != C - 1 || last_bytes == kChunkBytes);
                    int nvec = full ? kChunkVecs : (last_bytes >> 4);
                    if (SKEW) {
                        full = cidx < b_last && (cidx != 0 || b_h == 0);  // interior chunk: whole vectors only
                        nvec = full ? kChunkVecs : min(max(b_nvec - cidx * kChunkVecs, 0), kChunkVecs);
                        if (!full && nvec == kChunkVecs && !(cidx == 0 && b_h != 0) && !(b_keep != 0 && cidx == b_last))
                            full = true;  // the last chunk happens to be complete and to end on a vector boundary
                    }
                    if (SKEW && DIRECT && !full) {
                        // a chunk holding the first or the last vector of the span (or a short one): the edge vectors
                        // leave element by element, everything else as whole vectors
                        uint4* gv = reinterpret_cast<uint4*>(grow + static_cast<int64_t>(cidx) * kChunkElems);
                        const int v0 = cidx * kChunkVecs;
                        auto put = [&](int v, const uint4& o) {
                            const int lo = (v0 + v == 0) ? b_h : 0;
                            const int hi = (v0 + v == b_nvec - 1 && b_keep != 0) ? b_keep : 8;
                            if (lo != 0 || hi != 8)
                                st_global_edge(gv + v, o, lo, hi);
                            else
                                st_global_cs(gv + v, o);
                        };
                        if (nvec == kChunkVecs && !zero_row) {
                            uint4 v[kVpt];
#pragma unroll
                            for (int k = 0; k < kVpt; ++k) v[k] = sv[tid + k * kConsumers];
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&sm.done_bar[bcur.slot]);
#pragma unroll
                            for (int k = 0; k < kVpt; ++k) put(tid + k * kConsumers, grad_vec<F16>(v[k], c2, nl2, ng2));
                        } else {
                            for (int v = tid; v < nvec; v += kConsumers)
                                put(v, zero_row ? make_uint4(0u, 0u, 0u, 0u) : grad_vec<F16>(sv[v], c2, nl2, ng2));
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&sm.done_bar[bcur.slot]);
                        }
                        if (!zero_row && patch_mine && cidx == rr.id_chunk)
                            el_store<F16>(gv, rr.id_elem, rr.patch);
                    } else if (DIRECT) {
                        // gradients go straight from registers to global memory (coalesced 16-byte stores); the slot is
                        // released as soon as every warp has READ it, so its refill overlaps this chunk's arithmetic
                        uint4* gv = reinterpret_cast<uint4*>(grow + static_cast<int64_t>(cidx) * kChunkElems);
                        if (zero_row) {
                            if (lane == 0) mbar_arrive(&sm.done_bar[bcur.slot]);
                            for (int v = tid; v < nvec; v += kConsumers) st_global_cs(gv + v, make_uint4(0u, 0u, 0u, 0u));
                        } else if (full) {
                            uint4 v[kVpt];
#pragma unroll
                            for (int k = 0; k < kVpt; ++k) v[k] = sv[tid + k * kConsumers];
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&sm.done_bar[bcur.slot]);
#pragma unroll
                            for (int k = 0; k < kVpt; ++k) st_global_cs(gv + tid + k * kConsumers, grad_vec<F16>(v[k], c2, nl2, ng2));
                        } else {
                            for (int v = tid; v < nvec; v += kConsumers) st_global_cs(gv + v, grad_vec<F16>(sv[v], c2, nl2, ng2));
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&sm.done_bar[bcur.slot]);
                        }
                        if (!zero_row && patch_mine && cidx == rr.id_chunk)  // same thread wrote the vector holding it
                            el_store<F16>(gv, rr.id_elem, rr.patch);
                    } else {
                        if (zero_row) {
                            for (int v = tid; v < nvec; v += kConsumers) sv[v] = make_uint4(0u, 0u, 0u, 0u);
                        } else {
                            if (full) {
                                uint4 v[kVpt];
#pragma unroll
                                for (int k = 0; k < kVpt; ++k) v[k] = sv[tid + k * kConsumers];
#pragma unroll
                                for (int k = 0; k < kVpt; ++k) sv[tid + k * kConsumers] = grad_vec<F16>(v[k], c2, nl2, ng2);
                            } else {
                                for (int v = tid; v < nvec; v += kConsumers) sv[v] = grad_vec<F16>(sv[v], c2, nl2, ng2);
                            }
                            if (patch_mine && cidx == rr.id_chunk)
                                el_store<F16>(sv, rr.id_elem, rr.patch);
                        }
                        fence_proxy_async();  // generic-proxy writes -> visible to the bulk store
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.done_bar[bcur.slot]);
                    }
                    tr.ev(8, i, cidx);
                    bcur.advance(num_slots);
                }
            }
        }
        tr.finish();
    }
#ifdef B200TRL_K1_TRACE
    __syncthreads();
    if (a.trace && blockIdx.x < kTraceCtas) {
        unsigned long long* out = a.trace + static_cast<size_t>(blockIdx.x) * kTraceRoles * (kTraceEvents + 1);
        for (int r = 0; r < kTraceRoles; ++r) {
            if (tid == 0) out[r * (kTraceEvents + 1)] = static_cast<unsigned long long>(sm.trace.n[r]);
            for (int k = tid; k < kTraceEvents; k += blockDim.x) out[r * (kTraceEvents + 1) + 1 + k] = sm.trace.ev[r][k];
        }
    }
#endif
    // no CTA may leave while a peer can still address its shared memory
    __syncwarp();
    if (csize > 1) cluster_sync_all();
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// extra_bytes: 16 when the rows are skewed (their aligned span is up to 14 bytes longer than the slice)
int pick_cluster(int64_t vocab, int num_slots, int chunk_bytes, int extra_bytes) {
    static const int forced = env_int("B200TRL_K1_CLUSTER", 0);  // tuning knob: force a (larger) cluster size
    if (forced == 3 || forced == 5 || forced == 6 || forced == 7) {  // experiments: any size the hardware takes
        const int64_t slice = ((vocab + forced - 1) / forced + 7) & ~int64_t(7);
        const int64_t chunks = (slice * 2 + extra_bytes + chunk_bytes - 1) / chunk_bytes;
        if (chunks <= num_slots - 1) return forced;
    }
    for (int cs = 1; cs <= kMaxCluster; cs *= 2) {
        if (cs < forced) continue;
        const int64_t slice = ((vocab + cs - 1) / cs + 7) & ~int64_t(7);
        const int64_t chunks = (slice * 2 + extra_bytes + chunk_bytes - 1) / chunk_bytes;
        if (chunks <= num_slots - 1 || (chunks <= num_slots && cs == kMaxCluster)) return cs;
    }
    return 0;
}

template <bool F, bool Bk, bool DUAL, int NC, bool DIRECT, bool SKIP, int FAST, bool SKEW = false, bool F16 = false>
int launch_mode_f(const K1Args& a, int cs, int num_slots, cudaStream_t stream) {
    auto kern = k1_resident_kernel<F, Bk, DUAL, NC, DIRECT, SKIP, FAST, SKEW, F16>;
    constexpr int kThreads = NC + 64;  // + DMA warp + reducer warp
    static_assert(chunk_bytes_for(NC) % (NC * 16) == 0, "a full chunk must give every consumer the same vector count");
    constexpr int kCtasPerSm = (NC <= 256) ? 2 : 1;
    const size_t smem = static_cast<size_t>(num_slots) * chunk_bytes_for(NC) + sizeof(Smem);
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) {
            set_error("k1_resident: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
            return B200TRL_E_LAUNCH;
        }
    }
    const int sms = num_sms();
    int64_t clusters = std::min<int64_t>(static_cast<int64_t>(sms) * kCtasPerSm / cs, a.n_rows);
    if (clusters < 1) clusters = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(clusters * cs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent kernel with a static row stride: every cluster of the grid must be co-resident, otherwise the
    // late ones form a second wave.  Ask the driver how many clusters of this shape fit (GPC geometry limits it).
    {
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0 &&
            max_clusters < clusters) {
            clusters = max_clusters;
            cfg.gridDim = dim3(static_cast<unsigned>(clusters * cs));
        } else {
            (void)cudaGetLastError();
        }
    }
    static const int cap_clusters = env_int("B200TRL_K1_MAXCLUSTERS", 0);  // tuning knob: use fewer clusters
    if (cap_clusters > 0 && cap_clusters < clusters) {
        clusters = cap_clusters;
        cfg.gridDim = dim3(static_cast<unsigned>(clusters * cs));
    }
    static const int max_lag = std::min(2, std::max(0, env_int("B200TRL_K1_LAG", 1)));  // 1 measured best (0: -10 %, 2: -3 %)
    static const int l2_prefetch = env_int("B200TRL_K1_L2PREFETCH", 0);
    static const int verbose = env_int("B200TRL_K1_VERBOSE", 0);
    if (verbose)
        fprintf(stderr, "k1_resident<fwd=%d,bwd=%d> consumers=%d cluster=%d slots=%d smem=%zu clusters=%lld\n", (int)F, (int)Bk,
                NC, cs, num_slots, smem, (long long)clusters);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, num_slots, max_lag, Bk ? l2_prefetch : 0);
    if (e != cudaSuccess) {
        set_error("k1_resident launch failed: %s", cudaGetErrorString(e));
        return B200TRL_E_LAUNCH;
    }
    return check_launch("k1_resident_kernel");
}

// The fused pass with direct stores and two chains runs on FusedConsumer (B200TRL_K1_FAST=0: generic consumer code).
bool rows_skewed(const K1Args& a) {
    return a.vocab % 8 != 0 || a.row_stride % 8 != 0 || a.batch_stride % 8 != 0 ||
           (reinterpret_cast<uintptr_t>(a.logits) & 15) != 0;
}

template <bool F, bool Bk, bool DUAL, int NC, bool DIRECT, bool SKIP>
int launch_mode_s(const K1Args& a, int cs, int num_slots, cudaStream_t stream) {
#if K1_UNIT_F16
    // fp16 logits: the generic consumer code with the fp16 unpack / pack / max (dual chains; lean shapes excluded)
    if constexpr (DUAL && NC != 320 && NC != 384) {
        if constexpr (DIRECT || !Bk) {
            if (rows_skewed(a)) return launch_mode_f<F, Bk, DUAL, NC, DIRECT, SKIP, 0, true, true>(a, cs, num_slots, stream);
        }
        if (!rows_skewed(a)) return launch_mode_f<F, Bk, DUAL, NC, DIRECT, SKIP, 0, false, true>(a, cs, num_slots, stream);
    }
    set_error("k1_resident: no fp16 instantiation for this variant");
    return B200TRL_E_UNSUPPORTED;
#else
    // skewed rows (vocab % 8 != 0 ...): generic consumers with masked edge vectors, dlogits straight from registers
    if constexpr (DUAL && (DIRECT || !Bk)) {
        if (rows_skewed(a)) return launch_mode_f<F, Bk, DUAL, NC, DIRECT, SKIP, 0, true>(a, cs, num_slots, stream);
    }
    if (rows_skewed(a)) {
        set_error("k1_resident: skewed rows need the dual-chain, direct-store variant");
        return B200TRL_E_UNSUPPORTED;
    }
#ifndef B200TRL_K1_TRACE
    if constexpr (F && Bk && DUAL && DIRECT) {
        static const int fast = env_int("B200TRL_K1_FAST", 1);
        if (fast != 0) return launch_mode_f<F, Bk, DUAL, NC, DIRECT, SKIP, 1>(a, cs, num_slots, stream);
    }
#endif
    return launch_mode_f<F, Bk, DUAL, NC, DIRECT, SKIP, 0>(a, cs, num_slots, stream);
#endif  // K1_UNIT_F16
}

// Four CTA geometries (which one serves a call: pick_geom below):
//   mid   : 640 consumers, 11 x 20 KB slots (220 KB), 1 CTA / SM  — the fused pass's usual shape
//   wide  : 512 consumers, 13 x 16 KB slots (208 KB), 1 CTA / SM  — a row slice of up to 12 chunks per CTA
//   dense : 768 consumers,  9 x 24 KB slots (216 KB), 1 CTA / SM  — backward-only streams through 4-6 of them
//   twin  : 256 consumers,  6 x 16 KB slots ( 96 KB), 2 CTAs / SM — forward-only at any vocabulary, fused for rows
//           <= 80 KB; the two CTAs of an SM drift apart, so one computes while the other waits on a barrier or on HBM
struct Geom {
    int cs, slots, nc;
};
constexpr int kTwinSlots = 6;
constexpr int kDenseSlots = 9;  // 9 x 24 KB = 216 KB
constexpr int kMidSlots = 11;   // 11 x 20 KB = 220 KB

enum Mode { M_FWD, M_BWD, M_FUSED };
Mode mode_of(const K1Args& a) {
    const bool fwd = (a.lse_in == nullptr), bwd = (a.dlogits != nullptr);
    return (fwd && bwd) ? M_FUSED : (fwd ? M_FWD : M_BWD);
}

Geom pick_geom(int64_t vocab, Mode m, bool skew = false) {
    static const int mode = env_int("B200TRL_K1_GEOM", 0);  // 0 auto, 1 wide, 2 twin, 3 dense, 4 mid
    const int extra = skew ? 16 : 0;
    Geom wide{pick_cluster(vocab, kMaxSlots, kChunkBytes, extra), kMaxSlots, 512};
    Geom twin{pick_cluster(vocab, kTwinSlots, kChunkBytes, extra), kTwinSlots, 256};
    Geom dense{pick_cluster(vocab, kDenseSlots, chunk_bytes_for(768), extra), kDenseSlots, 768};
    const Geom mid{pick_cluster(vocab, kMidSlots, chunk_bytes_for(640), extra), kMidSlots, 640};
    if (m != M_FUSED) {
        // forward-only / backward-only: nothing has to stay resident between two sweeps, so a row simply STREAMS
        // through the ring of one CTA -- no cluster, no DSMEM exchange, any vocabulary.  Measured at config 2
        // (tools/k1_sweep.sh): forward-only 0.85 ms (2-CTA cluster, dense) -> 0.76 ms = 100 % of the copy peak with
        // two independent 256-consumer CTAs per SM; backward-only 1.69 -> 1.61 ms with one dense CTA (-> 1.55 ms with
        // the short ring below).
        static const int fwd_cs = env_int("B200TRL_K1_FWD_CS", 1), bwd_cs = env_int("B200TRL_K1_BWD_CS", 1);
        const int cs = std::max(1, m == M_FWD ? fwd_cs : bwd_cs);
        wide.cs = std::min(wide.cs ? wide.cs : cs, cs);
        twin.cs = std::min(twin.cs ? twin.cs : cs, cs);
        dense.cs = std::min(dense.cs ? dense.cs : cs, cs);
        if (mode == 1) return wide;
        if (mode == 2) return twin;
        if (mode == 3) return dense;
        // backward-only reads a row and writes it back a ring's depth later: a SHORT ring (96-120 KB in flight per SM)
        // keeps the read and write streams of an SM close together in time and is both faster and insensitive to how
        // the two buffers are placed -- 9 / 5 / 4 slots: V = 151 936 1.71 / 1.60 / 1.55 ms, 128 256 1.42 / 1.34 / 1.30,
        // 65 536 1.44 / 1.34 / 1.37, 32 000 1.43 / 1.42 / 1.43 (tools/k1_sweep.sh, tools/k1_offset_probe.py)
        static const int bwd_slots = env_int("B200TRL_K1_BWD_SLOTS", 0);
        if (m == M_BWD) {
            const int64_t row_bytes = vocab * 2, chunk = chunk_bytes_for(768);
            const int64_t chunks = (row_bytes + chunk - 1) / chunk, tail = row_bytes % chunk;
            int slots = row_bytes >= 200000 ? 4 : 5;
            // a row that fills the ring exactly and ends in a sliver of a chunk (V = 50 304: 4 x 24 KB + 2.3 KB) would
            // serialise on that sliver's slot: one more slot (measured 1.10 -> 1.00 ms; 57 344 and 49 152 prefer 5)
            if (chunks <= slots && tail != 0 && tail < chunk / 4) slots = static_cast<int>(chunks) + 1;
            dense.slots = bwd_slots >= 2 && bwd_slots <= kDenseSlots ? bwd_slots : slots;
        }
        return m == M_FWD ? twin : dense;
    }
    if (mode == 4 && mid.cs) return mid;
    if (mode == 5 && mid.cs) return Geom{mid.cs, kMidSlots, 320};    // lean mid: half the warps, twice the work per wait
    if (mode == 6 && dense.cs) return Geom{dense.cs, kDenseSlots, 384};
    if (mode == 3 && dense.cs && dense.cs <= wide.cs) return dense;
    if (mode == 2 && twin.cs) return twin;
    if (mode == 1) return wide;
    // measured (tools/k1_sweep.sh, B200): rows that fit one twin CTA (<= 80 KB, e.g. V = 32000) are fastest with two
    // drifting CTAs per SM (fused 91.7 % of the measured HBM peak); anything that would need a cluster in twin form is
    // faster with 1 CTA / SM
    if (twin.cs == 1) return twin;
    // 1 CTA / SM.  What separates the three shapes is (a) the cluster they need and (b) the time lost in a slice's last,
    // partial chunk: every chunk is a barrier round, and a partial one costs one or two vector-times whatever it holds.
    // Take the smallest cluster, then the least tail waste, then (within 1.5 %) 640 > 768 > 512 consumers.  Measured,
    // fused, % of the HBM peak for 512 / 640 / 768: V = 151 936 90.3 / 93.1 / 88.9, 128 256 89.6 / 89.7 / 87.4, 100 352
    // 82.4 / 94.5 / 82.9 (only 640 fits one CTA), 65 536 92.0 / 92.1 / 90.9, 50 304 82.0 / 89.9 / 82.7, 49 152 88.9 /
    // 88.8 / 90.4.
    auto waste = [&](const Geom& g) {
        const int chunk = chunk_bytes_for(g.nc);
        const int64_t slice_vecs = (((vocab + g.cs - 1) / g.cs + 7) & ~int64_t(7)) / 8;
        const int64_t chunk_vecs = chunk / 16, full = slice_vecs / chunk_vecs, rest = slice_vecs - full * chunk_vecs;
        const double vpt = static_cast<double>(chunk_vecs) / g.nc;  // vector-times of a full round
        const double rounds = static_cast<double>(full) + (rest ? static_cast<double>((rest + g.nc - 1) / g.nc) / vpt : 0.0);
        return rounds * static_cast<double>(chunk_vecs) / static_cast<double>(slice_vecs) - 1.0;
    };
    const Geom* order[3] = {&mid, &dense, &wide};
    const Geom* best = nullptr;
    for (const Geom* g : order) {
        if (!g->cs) continue;
        if (!best || g->cs < best->cs || (g->cs == best->cs && waste(*g) < waste(*best) - 0.015)) best = g;
    }
    Geom out = best ? *best : wide;
    static const int fused_slots = env_int("B200TRL_K1_FUSED_SLOTS", 0);  // tuning knob: shorter ring (>= chunks + 1)
    if (fused_slots > 0 && fused_slots < out.slots) {
        const int64_t slice = ((vocab + out.cs - 1) / out.cs + 7) & ~int64_t(7);
        const int chunks = static_cast<int>((slice * 2 + chunk_bytes_for(out.nc) - 1) / chunk_bytes_for(out.nc));
        out.slots = std::max(fused_slots, chunks + 1);
    }
    return out;
}

template <bool F, bool Bk, bool DUAL, int NC, bool DIRECT>
int launch_mode_p(const K1Args& a, int cs, int num_slots, cudaStream_t stream) {
    constexpr bool kFused = F && Bk;
    if (kFused && a.skip_masked && (a.gmode == G_GRPO || a.gmode == G_PPO))
        return launch_mode_s<F, Bk, DUAL, NC, DIRECT, kFused>(a, cs, num_slots, stream);
    // backward-only with a per-token upstream gradient: rows whose gradient is exactly zero (masked tokens) are not
    // read at all -- their dlogits are zeros whatever the logits hold.  Always on: no output changes.
    constexpr bool kFwdOnly = F && !Bk;
    if (kFwdOnly && a.row_mask) return launch_mode_s<F, Bk, DUAL, NC, DIRECT, kFwdOnly>(a, cs, num_slots, stream);
    constexpr bool kBwdOnly = !F && Bk;
    static const int skip_zero = env_int("B200TRL_K1_SKIPZERO", 1);
    if (kBwdOnly && skip_zero && a.gmode == G_GIVEN)
        return launch_mode_s<F, Bk, DUAL, NC, DIRECT, kBwdOnly>(a, cs, num_slots, stream);
    return launch_mode_s<F, Bk, DUAL, NC, DIRECT, false>(a, cs, num_slots, stream);
}

// DIRECT: the consumers store dlogits themselves (registers -> global) instead of rewriting the slot for a TMA bulk
// store; the slot is then free as soon as it has been read.  Env knob B200TRL_K1_DIRECT: -1 auto, 0 bulk stores, 1 direct.
template <bool F, bool Bk, bool DUAL, int NC>
int launch_mode_t(const K1Args& a, int cs, int num_slots, cudaStream_t stream) {
    if (!Bk) return launch_mode_p<F, Bk, DUAL, NC, false>(a, cs, num_slots, stream);
    static const int direct_env = env_int("B200TRL_K1_DIRECT", -1);
    const bool direct = rows_skewed(a) || (direct_env < 0 ? (F && Bk) : direct_env != 0);
    return direct ? launch_mode_p<F, Bk, DUAL, NC, Bk>(a, cs, num_slots, stream)
                  : launch_mode_p<F, Bk, DUAL, NC, false>(a, cs, num_slots, stream);
}

template <bool F, bool Bk>
int launch_mode(const K1Args& a, const Geom& g, cudaStream_t stream) {
    // Two accumulation chains with one reference-point check per vector pair: measured on B200 (tools/k1_sweep.sh)
    // they help every mode (forward-only +3 %, fused +2..4 %), so the single-chain form is no longer instantiated
    // (the DUAL template parameter stays for the record; it halves the number of kernels to compile).
    if (g.nc == 256) return launch_mode_t<F, Bk, true, 256>(a, g.cs, g.slots, stream);
    if (g.nc == 640)  // fused pass only (pick_geom); the other modes never ask for it
        return launch_mode_t<F, Bk, true, (F && Bk) ? 640 : 512>(a, g.cs, g.slots, stream);
    if (g.nc == 768) return launch_mode_t<F, Bk, true, 768>(a, g.cs, g.slots, stream);
    if constexpr (F && Bk) {  // lean shapes: fused pass only
        if (g.nc == 320) return launch_mode_t<F, Bk, true, 320>(a, g.cs, g.slots, stream);
        if (g.nc == 384) return launch_mode_t<F, Bk, true, 384>(a, g.cs, g.slots, stream);
    }
    return launch_mode_t<F, Bk, true, 512>(a, g.cs, g.slots, stream);
}

}  // namespace

#if !K1_UNIT_F16
bool k1_resident_supported(const K1Args& a, int dtype) {
    if (dtype != B200TRL_BF16 && dtype != B200TRL_F16) return false;
    static const int allow_f16 = env_int("B200TRL_K1_F16", 1);  // 0: fp16 logits go to the row kernel (A/B runs)
    if (dtype == B200TRL_F16 && !allow_f16) return false;
    static const int allow_skew = env_int("B200TRL_K1_SKEW", 1);  // 0: skewed rows go to the row kernel (A/B runs)
    const bool skew = rows_skewed(a);
    if (skew && !allow_skew) return false;
    if ((reinterpret_cast<uintptr_t>(a.logits) & 1) != 0) return false;
    if (a.dlogits) {
        // every dlogits row must sit at the same offset inside its 16-byte granule as the logits row it mirrors
        // (true for `empty_like(logits)` and for the in-place form): the interior vectors are then aligned stores
        const uintptr_t la = reinterpret_cast<uintptr_t>(a.logits), da = reinterpret_cast<uintptr_t>(a.dlogits);
        if (((la ^ da) & 15) != 0 || (a.dl_row_stride - a.row_stride) % 8 != 0) return false;
        if (a.rows_per_batch != 0 && (a.dl_batch_stride - a.batch_stride) % 8 != 0) return false;
    }
    if (a.vocab * 2 < 2 * kChunkBytes) return false;  // tiny rows: per-row overheads dominate, use the row kernel
    if (a.vocab * 2 > (int64_t(1) << 30)) return false;  // slice bytes are kept in 32-bit counters
    return pick_geom(a.vocab, mode_of(a), skew).cs != 0;
}

// AUTO policy: whenever this kernel can take the call it is the faster one.  Measured fused / forward-only against
// the row kernel (tools/k1_sweep.sh): V = 200 000 (4-CTA cluster) 1.36 / 0.51 ms vs 1.69 / 0.79 ms; V = 262 144
// 1.53 / 0.65 vs 2.28 / 0.99; V = 524 288 (8-CTA cluster) 1.66 / 0.65 vs 2.32 / 0.92.
bool k1_resident_preferred(const K1Args& a, int dtype) { return k1_resident_supported(a, dtype); }

// host-only query behind b200trl_k1_geometry: {consumer threads, cluster size, ring slots, chunk bytes}, 0s if the
// resident kernel cannot take the vocabulary in that mode (0 forward-only, 1 backward-only, 2 fused)
void k1_resident_geometry(int64_t vocab, int mode, int32_t out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0;
    if (vocab <= 0 || vocab * 2 < 2 * kChunkBytes || mode < 0 || mode > 2) return;
    const Geom g = pick_geom(vocab, mode == 0 ? M_FWD : (mode == 1 ? M_BWD : M_FUSED), vocab % 8 != 0);
    if (!g.cs) return;
    out[0] = g.nc;
    out[1] = g.cs;
    out[2] = g.slots;
    out[3] = chunk_bytes_for(g.nc);
}

#endif  // !K1_UNIT_F16

#if K1_UNIT_F16
int launch_k1_resident_f16(const K1Args& a, cudaStream_t stream) {
#else
int launch_k1_resident_f16(const K1Args& a, cudaStream_t stream);  // k1_resident_f16.cu
int launch_k1_resident(const K1Args& a, cudaStream_t stream) {
    if (a.elem_f16) return launch_k1_resident_f16(a, stream);
#endif
    if (a.n_rows == 0) return B200TRL_OK;
    const Mode m = mode_of(a);
    const Geom g = pick_geom(a.vocab, m, rows_skewed(a));
    B200TRL_REQUIRE(g.cs != 0, B200TRL_E_UNSUPPORTED, "k1_resident: vocab %lld too large for an 8-CTA cluster",
                    (long long)a.vocab);
    if (m == M_FUSED) return launch_mode<true, true>(a, g, stream);
    if (m == M_FWD) return launch_mode<true, false>(a, g, stream);
    B200TRL_REQUIRE(a.dlogits != nullptr, B200TRL_E_INVALID, "k1_resident: nothing to do");
    return launch_mode<false, true>(a, g, stream);
}

}  // namespace b200trl
