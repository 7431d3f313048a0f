// Batched-affine bucket accumulation for the Pippenger MSM (the dominant kernel of the path).
//
// The XYZZ mixed addition of k_msm_accumulate costs 9.5 Fq products (2 736 IMAD.WIDE) and the kernel is bound by the
// integer multiply pipe.  An AFFINE addition whose inversion is shared with millions of others (Montgomery's trick)
// costs 6 products + ~0.4 for the shared product tree (1 843 IMAD.WIDE): this file sums the sorted entry stream with
// affine additions as far as that pays, then hands what is left to the XYZZ path.
//
// Shape.  The stream (sorted by bucket) is viewed as a complete binary tree over ALIGNED positions: block (m, i) =
// entries [i, i + 2^m) with i a multiple of 2^m.  A block is PURE when all its entries belong to one bucket (the stream
// is sorted, so: first key == last key).  Level k = 0 .. K-1 merges the two halves of every pure block of size 2^(k+1)
// -- by induction each half already is a single affine point -- with one device-wide batch inversion per level:
//     lvl      one pass: lvl[i] = size exponent of the largest pure aligned block that starts at i (<= K)
//     phase 1  denominator d = x2 - x1 of every pure pair (2 y1 for a doubling, 1 when no inverse is needed),
//              per-thread running products to scratch, thread totals to the product tree
//     tree     fan-in-32 prefix products up, ONE field inversion at the top, inverses of the thread totals down
//              (batch_inv.cuh)
//     phase 2  walks each thread's pairs backwards (1 / d_j = I * prefix_{j-1}, I *= d_j), slope, sum -> R[k+1]
// No compaction, no per-level scan, fixed index arithmetic: a bucket of 58 entries decomposes into ~5 maximal pure
// blocks; the ~52 additions inside them are affine, the ~5 between them (and everything a chunk boundary cuts) go
// through the XYZZ accumulate pass below, which steps through the stream block by block instead of entry by entry and
// keeps the contract of k_msm_accumulate (whole buckets -> buckets[], pieces -> parts[], fixed up by k_msm_fixup*).
// Every exceptional case (identity operands, P + P, P + (-P)) is handled by g1a_batch_denominator's case analysis.
// The stream is processed in slabs so that the level arrays (96 B per entry) stay inside a memory budget.
#include <stdlib.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "batch_inv.cuh"
#include "g1_batch_affine.cuh"
#include "msm.h"

namespace scz {

constexpr int BA_THREADS = 128;
#ifndef BA_B_PAIRS
#define BA_B_PAIRS 32   // measured per 2^20 proof: 8 -> 122.2, 16 -> 120.5, 32 -> 119.7 ms of accumulation (the product tree shrinks)
#endif
constexpr int BA_B = BA_B_PAIRS;   // pairs per thread in phase 1 / 2 (strided by the CTA: coalesced across lanes)
constexpr int BA_MAX_LEVELS = 8;
constexpr int BA_ACC_THREADS = 128;

struct BaLevels {
    void *R[BA_MAX_LEVELS + 1];   // R[m]: one affine point per aligned 2^m-block of the slab (R[0] unused: the bases)
};

// bases pointer of the segment that owns bucket `key`, cached while consecutive keys stay inside one segment
struct SegCursor {
    uint32_t lo = 1, hi = 0;
    const void *bases = nullptr;
    __device__ __forceinline__ const void *get(const MsmSeg *segs, int nseg, uint32_t key) {
        if (key < lo || key >= hi) {
            int s = nseg == 1 ? 0 : seg_by_bucket(segs, nseg, key);
            lo = __ldg(&segs[s].bucket_base);
            hi = lo + __ldg(&segs[s].W) * __ldg(&segs[s].nb);
            bases = segs[s].bases;
        }
        return bases;
    }
};

// ---- lvl[j] for position i = slab_base + j: the largest m <= K with i % 2^m == 0, i + 2^m <= E and
//      key[i] == key[i + 2^m - 1].  slab_base is a multiple of 2^K, so alignment can be read off j.
__global__ void __launch_bounds__(256) k_ba_levels(const uint32_t *E_ptr, const uint2 *sorted, uint32_t slab_base,
                                                   uint32_t slab_len, uint32_t K, uint8_t *lvl) {
    uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= slab_len) return;
    const uint32_t E = __ldg(E_ptr);
    uint64_t i = (uint64_t)slab_base + j;
    uint8_t m = 0;
    if (i < E) {
        uint32_t key = __ldg(&sorted[i].y);
        uint32_t maxm = j ? min(K, (uint32_t)(__ffs(j) - 1)) : K;
        for (uint32_t mm = 1; mm <= maxm; mm++) {
            uint64_t last = i + (1u << mm) - 1;
            if (last >= E || __ldg(&sorted[last].y) != key) break;
            m = (uint8_t)mm;
        }
    }
    lvl[j] = m;
}

// the point that stands for block (k, rel) of the slab: level 0 = the entry's base (negated when the digit is negative)
template <bool L0>
__device__ __forceinline__ G1Affine ba_point(const MsmSeg *segs, int nseg, SegCursor &sc, const uint2 *sorted,
                                             uint32_t slab_base, uint32_t k, uint32_t rel, const void *Rk) {
    if (L0) {
        uint2 e = __ldg(sorted + (size_t)slab_base + rel);
        G1Affine p = g1a_load_stream(sc.get(segs, nseg, e.y), e.x & 0x7fffffffu);
        if (e.x >> 31) p.y = fp_neg(p.y);
        return p;
    }
    G1Affine p;
    const char *b = reinterpret_cast<const char *>(Rk) + (size_t)(rel >> k) * 96;
    p.x = fp_load_rw<FqP>(b, 0);
    p.y = fp_load_rw<FqP>(b, 1);
    return p;
}
template <bool L0>
__device__ __forceinline__ Fq ba_point_x(const MsmSeg *segs, int nseg, SegCursor &sc, const uint2 *sorted, uint32_t slab_base,
                                         uint32_t k, uint32_t rel, const void *Rk) {
    if (L0) {
        uint2 e = __ldg(sorted + (size_t)slab_base + rel);
        const char *b = reinterpret_cast<const char *>(sc.get(segs, nseg, e.y)) + (size_t)(e.x & 0x7fffffffu) * 96;
        Fq x;
        const uint4 *q = reinterpret_cast<const uint4 *>(b);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            uint4 v = __ldcg(q + i);
            x.l[4 * i] = v.x, x.l[4 * i + 1] = v.y, x.l[4 * i + 2] = v.z, x.l[4 * i + 3] = v.w;
        }
        return x;
    }
    return fp_load_rw<FqP>(reinterpret_cast<const char *>(Rk) + (size_t)(rel >> k) * 96, 0);
}

__device__ __noinline__ Fq ba_denominator_slow(const G1Affine &p, const G1Affine &q) {
    BatchAddCase kind;
    return g1a_batch_denominator(p, q, kind);
}

// one (slab, level) work item of the affine tree
struct BaItem {
    const uint8_t *lvl;
    uint32_t slab_base, k, npairs;
    const void *Rk;       // level-k points of the slab (null at level 0: the bases)
    void *pre;            // scratch: per-thread running products
    void *tot;            // phase 1 out: thread totals (= V[0] of the product tree)
    const void *inv_tot;  // phase 2 in: their inverses (= P[0] after the tree)
    void *Rk1;            // phase 2 out: level-(k+1) points
};

// ---- phase 1 of level k: running products of the denominators.  Pair p of the slab = the two halves of block
//      (k + 1, p << (k + 1)).  pre[p] = product of this thread's denominators up to and including pair p.
template <bool L0>
__device__ __forceinline__ void ba_phase1_cta(const MsmSeg *segs, int nseg, const uint2 *sorted, const BaItem &it, uint32_t cta) {
    const uint32_t cta_base = cta * (BA_THREADS * BA_B), k = it.k;
    Fq run = Fq::one();
    SegCursor sc;
    for (int j = 0; j < BA_B; j++) {
        uint32_t p = cta_base + j * BA_THREADS + threadIdx.x;
        if (p >= it.npairs) break;
        uint32_t rel = p << (k + 1);
        if (it.lvl[rel] > k) {
            Fq x1 = ba_point_x<L0>(segs, nseg, sc, sorted, it.slab_base, k, rel, it.Rk);
            Fq x2 = ba_point_x<L0>(segs, nseg, sc, sorted, it.slab_base, k, rel + (1u << k), it.Rk);
            Fq d = fp_sub(x2, x1);
            if (d.is_zero() || x1.is_zero() || x2.is_zero()) {   // rare: equal / opposite points, a possible identity
                G1Affine P = ba_point<L0>(segs, nseg, sc, sorted, it.slab_base, k, rel, it.Rk);
                G1Affine Q = ba_point<L0>(segs, nseg, sc, sorted, it.slab_base, k, rel + (1u << k), it.Rk);
                d = ba_denominator_slow(P, Q);
            }
            run = fp_mul(run, d);
        }
        fp_store<FqP>(it.pre, p, run);
    }
    fp_store<FqP>(it.tot, (size_t)cta * BA_THREADS + threadIdx.x, run);
}

// the cases that are not "two distinct finite points with different x"
__device__ __noinline__ void ba_add_slow(const G1Affine &p, const G1Affine &q, const Fq &dinv, G1Affine &r) {
    BatchAddCase kind;
    (void)g1a_batch_denominator(p, q, kind);
    if (kind == BA_TAKE_P) r = p;
    else if (kind == BA_TAKE_Q) r = q;
    else if (kind == BA_INF) {
        r.x = Fq::zero();
        r.y = Fq::zero();
    } else {
        Fq num;
        if (kind == BA_ADD) num = fp_sub(q.y, p.y);
        else {   // tangent: 3 x^2 / 2 y
            Fq xx = fp_sqr(p.x);
            num = fp_add(fp_dbl(xx), xx);
        }
        Fq lam = fp_mul(num, dinv);
        r.x = fp_sub(fp_sub(fp_sqr(lam), p.x), q.x);
        r.y = fp_sub(fp_mul(lam, fp_sub(p.x, r.x)), p.y);
    }
}

// ---- phase 2 of level k: inv_tot[thread] = inverse of the thread's total (from the product tree); results to R[k+1]
template <bool L0>
__device__ __forceinline__ void ba_phase2_cta(const MsmSeg *segs, int nseg, const uint2 *sorted, const BaItem &it, uint32_t cta) {
    const uint32_t cta_base = cta * (BA_THREADS * BA_B), k = it.k;
    Fq I = fp_load_rw<FqP>(it.inv_tot, (size_t)cta * BA_THREADS + threadIdx.x);
    SegCursor sc;
    for (int j = BA_B - 1; j >= 0; j--) {
        uint32_t p = cta_base + j * BA_THREADS + threadIdx.x;
        if (p >= it.npairs) continue;
        uint32_t rel = p << (k + 1);
        if (it.lvl[rel] <= k) continue;
        G1Affine P = ba_point<L0>(segs, nseg, sc, sorted, it.slab_base, k, rel, it.Rk);
        G1Affine Q = ba_point<L0>(segs, nseg, sc, sorted, it.slab_base, k, rel + (1u << k), it.Rk);
        Fq dinv = j ? fp_mul(I, fp_load_rw<FqP>(it.pre, p - BA_THREADS)) : I;   // 1 / d_j
        Fq d = fp_sub(Q.x, P.x);
        G1Affine r;
        if (d.is_zero() || P.x.is_zero() || Q.x.is_zero()) {
            // copies, so that only they have their address taken: with P, Q, dinv and r themselves passed by reference
            // to the out-of-line slow path every pair stored ~250 B to the stack frame before this (rare) branch
            const G1Affine Pc = P, Qc = Q;
            const Fq dc = dinv;
            G1Affine rc;
            d = ba_denominator_slow(Pc, Qc);
            ba_add_slow(Pc, Qc, dc, rc);
            r = rc;
        } else {
            Fq lam = fp_mul(fp_sub(Q.y, P.y), dinv);
            r.x = fp_sub(fp_sub(fp_sqr(lam), P.x), Q.x);   // (a dedicated squaring, fq_sqr_sos, measured slower: +1.5 ms per proof)
            r.y = fp_sub(fp_mul(lam, fp_sub(P.x, r.x)), P.y);
        }
        if (j) I = fp_mul(I, d);                                                // 1 / (d_0 ... d_{j-1})
        g1a_store(it.Rk1, p, r);
    }
}

// (measured, not kept: level-0 phase 1 with two pairs per step -- one 16-byte load for the two adjacent entries of a pair,
// the four x gathers of a step issued before the first product, 88 registers: 120.4 vs 119.8 ms of accumulation per 2^20
// proof; the kernel's DRAM rate is set by the random 128-byte line fills, not by the loads in flight.)
// (measured, not kept: level-0 phase 2 with its two gathered operands staged per thread in shared memory by cp.async,
// two pairs deep, four CTAs per SM -- 121.0 instead of 119.5 ms of accumulation per 2^20 proof.  The level-0 kernels are
// bound by the DRAM rate of random 96 B gathers (ncu: 2.6 - 3.9 TB/s of 128 B line fills), not by exposed latency.)
__global__ void __launch_bounds__(BA_THREADS) k_ba_phase1(const MsmSeg *segs, int nseg, const uint2 *sorted, BaItem it) {
    if (it.k == 0) ba_phase1_cta<true>(segs, nseg, sorted, it, blockIdx.x);
    else ba_phase1_cta<false>(segs, nseg, sorted, it, blockIdx.x);
}
#ifndef BA_P2_MIN_BLOCKS
#define BA_P2_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(BA_THREADS, BA_P2_MIN_BLOCKS) k_ba_phase2(const MsmSeg *segs, int nseg, const uint2 *sorted, BaItem it) {
    if (it.k == 0) ba_phase2_cta<true>(segs, nseg, sorted, it, blockIdx.x);
    else ba_phase2_cta<false>(segs, nseg, sorted, it, blockIdx.x);
}
// (measured, not kept: phase 2 of one slab half and phase 1 of the other in ONE grid with the two kinds of CTA interleaved,
// so that the memory-bound and the pipe-bound phase share every SM: 134 instead of 127 ms -- phase 2 needs its occupancy)

// ---- the XYZZ pass over what the affine levels left: same chunks, same outputs as k_msm_accumulate, but the walk
//      steps from one maximal pure block to the next (lvl[] gives its size, R[m] its sum) instead of entry by entry.
//      The loop is split: a light phase in which every lane walks on its own (run boundaries, first blocks of a run,
//      identity blocks) until it holds an addition that must really be done, then ONE mixed addition for all lanes of
//      the warp that have one -- the expensive code always runs with as many lanes as there is work.
#ifndef BA_ACC_MIN_BLOCKS
#define BA_ACC_MIN_BLOCKS 3   // 170 registers: three CTAs per SM hide the walk's dependent loads (measured: -1.4 ms per 2^20 proof)
#endif
__global__ void __launch_bounds__(BA_ACC_THREADS, BA_ACC_MIN_BLOCKS) k_ba_accumulate(const MsmSeg *segs, int nseg, const uint32_t *E_ptr,
                                                                   uint32_t logT, const uint2 *sorted, const uint32_t *counts,
                                                                   const uint32_t *cursor, const uint8_t *lvl,
                                                                   uint32_t slab_base, uint32_t slab_len, BaLevels L,
                                                                   void *buckets, void *parts) {
    const uint32_t E = __ldg(E_ptr);
    const uint32_t tl = blockIdx.x * BA_ACC_THREADS + threadIdx.x;      // chunk of the slab
    const uint64_t lo64 = (uint64_t)slab_base + ((uint64_t)tl << logT);
    bool done = lo64 >= E || ((uint64_t)tl << logT) >= slab_len;        // lanes without a chunk idle through the votes
    uint32_t t = 0, i = 0, hi = 0, cur = 0;
    bool head_piece = false, tail_piece = false, first_run = true;
    if (!done) {
        t = (uint32_t)(lo64 >> logT);                                   // global chunk index
        i = (uint32_t)lo64;
        hi = lo64 + (1u << logT) < E ? i + (1u << logT) : E;
        const uint32_t k_first = __ldg(&sorted[i].y), k_last = __ldg(&sorted[hi - 1].y);
        head_piece = cursor[k_first] - counts[k_first] < i;   // the first bucket began in an earlier chunk
        tail_piece = cursor[k_last] > hi;                     // the last bucket goes on in a later chunk
        cur = k_first;
    }
    SegCursor sc;
    G1X acc = G1X::inf();
    G1Affine p;
    p.x = Fq::zero();
    p.y = Fq::zero();
    auto flush = [&](bool last_run) {                         // the run of bucket `cur` ends
        if (first_run && head_piece) g1x_store(parts, 2 * (size_t)t, acc);
        else if (last_run && tail_piece) g1x_store(parts, 2 * (size_t)t + 1, acc);
        else g1x_store(buckets, cur, acc);
        first_run = false;
        acc = G1X::inf();
    };
    while (true) {
        bool pending = false;
        while (!done && !pending) {
            if (i >= hi) {
                flush(true);
                done = true;
                break;
            }
            const uint2 e = __ldg(sorted + i);
            const uint32_t m = lvl[i - slab_base];
            if (e.y != cur) {
                flush(false);
                cur = e.y;
            }
            if (m == 0) {
                p = g1a_load_stream(sc.get(segs, nseg, e.y), e.x & 0x7fffffffu);
                if (e.x >> 31) p.y = fp_neg(p.y);
            } else {
                const char *b = reinterpret_cast<const char *>(L.R[m]) + (size_t)((i - slab_base) >> m) * 96;
                p.x = fp_load_rw<FqP>(b, 0);
                p.y = fp_load_rw<FqP>(b, 1);
            }
            i += 1u << m;
            if (p.is_inf()) continue;
            if (acc.is_inf()) {
                acc.x = p.x, acc.y = p.y;
                acc.zz = Fq::one(), acc.zzz = Fq::one();
                continue;
            }
            pending = true;
        }
        if (!__any_sync(0xffffffffu, pending)) break;
        if (pending) g1x_add_affine(acc, p.x, p.y);
    }
}

// ------------------------------------------------------------------ host side
// The slab (entries worked on per pass) is sized from the DEVICE's memory and the number of ctxs that share the device,
// never from what happens to be free: the workspace is allocated once per ctx and kept.
static size_t ba_budget_bytes(Ctx *ctx) {
    static size_t total = [] {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
            cudaGetLastError();
            total_b = (size_t)32 << 30;
        }
        return total_b;
    }();
    const int live = std::max(1, ctx_live_on_device(ctx->device));
    return std::min((size_t)((double)total * 0.45 / live), (size_t)64 << 30);
}

namespace {
// the arrays of one slab: its levels, scratch and product tree
struct BaHalf {
    uint32_t base = 0, len = 0;
    uint64_t cap = 0;
    uint32_t levels = 0;
    std::unique_ptr<DevTmp> lvl, pre, R[BA_MAX_LEVELS + 1];
    InvTree<FqP> tree;   // values = the thread totals of phase 1, inverses = what phase 2 starts from (batch_inv.cuh)
    BaLevels L;
    int32_t alloc(Ctx *ctx, uint64_t cap_, uint32_t levels_) {
        cap = cap_, levels = levels_;
        lvl.reset(new DevTmp(ctx, true));
        pre.reset(new DevTmp(ctx, true));
        SCZ_TRY(lvl->alloc(cap));
        SCZ_TRY(pre->alloc((cap >> 1) * sizeof(Fq)));
        for (uint32_t m = 0; m <= BA_MAX_LEVELS; m++) {
            L.R[m] = nullptr;
            if (m >= 1 && m <= levels) {
                R[m].reset(new DevTmp(ctx, true));
                SCZ_TRY(R[m]->alloc((cap >> m) * sizeof(G1Affine)));
                L.R[m] = R[m]->p;
            }
        }
        SCZ_TRY(tree.alloc(ctx, (((cap >> 1) + BA_THREADS * BA_B - 1) / (BA_THREADS * BA_B)) * BA_THREADS, true));
        return SCZ_OK;
    }
    BaItem item(uint32_t k) const {
        BaItem it;
        it.lvl = lvl->as<uint8_t>();
        it.slab_base = base, it.k = k, it.npairs = len >> (k + 1);
        it.Rk = k ? L.R[k] : nullptr;
        it.pre = pre->p, it.tot = tree.values(), it.inv_tot = tree.inverses(), it.Rk1 = L.R[k + 1];
        return it;
    }
};
uint32_t ba_ctas(const BaItem &it) { return ceil_div_u32(it.npairs, BA_THREADS * BA_B); }

}   // namespace

int32_t msm_accumulate_affine(Ctx *ctx, const MsmSeg *sp, int nseg, const uint32_t *E_ptr, uint64_t entries, uint32_t logT,
                              const uint2 *sorted, const uint32_t *counts, const uint32_t *cursor, void *buckets, void *parts,
                              uint32_t levels) {
    if (levels < 1 || levels > BA_MAX_LEVELS || levels > logT) return ctx->fail(SCZ_ERR_BAD_ARG, "msm affine: bad level count %u", levels);
    cudaStream_t st = ctx->stream;
    const uint64_t T = 1ull << logT;
    const uint64_t entries_up = (entries + T - 1) / T * T;
    // bytes per slab entry: levels (1) + R[1..K] (<= 96) + scratch prefixes (24) + tree (~2 * 48 / BA_B = 12)
    const uint64_t per_entry = 1 + 96 + 24 + 12;
    uint64_t slab = ctx->msm_affine_slab ? ctx->msm_affine_slab : ba_budget_bytes(ctx) / per_entry;
    slab = std::max<uint64_t>(T, slab / T * T);
    slab = std::min<uint64_t>(slab, entries_up);
    if (slab >= (1ull << 32)) slab = (1ull << 32) - T;
    // as few slabs as the budget allows, all of the same length
    {
        const uint64_t nslabs = (entries_up + slab - 1) / slab;
        slab = ((entries_up + nslabs - 1) / nslabs + T - 1) / T * T;
    }
    // the ctx's workspace: kept across sequences, replaced only when a bigger one is needed
    if (slab > ctx->msm_affine_ws_limit) {
        const uint64_t lim = std::max<uint64_t>(T, ctx->msm_affine_ws_limit / T * T);
        const uint64_t nslabs = (entries_up + lim - 1) / lim;
        slab = std::min(lim, ((entries_up + nslabs - 1) / nslabs + T - 1) / T * T);
    }
    BaHalf *half = static_cast<BaHalf *>(ctx->msm_affine_ws.get());
    if (!half || half->levels < levels || half->cap < slab) {
        SCZ_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->msm_affine_ws.reset();
        while (true) {
            std::shared_ptr<BaHalf> ws(new BaHalf());
            if (ws->alloc(ctx, slab, std::max(levels, 4u)) == SCZ_OK) {
                ctx->msm_affine_ws = ws;
                half = ws.get();
                break;
            }
            ws.reset();
            if (slab <= (1ull << 20)) return ctx->fail(SCZ_ERR_NOMEM, "msm affine: no memory for a slab of %llu entries", (unsigned long long)slab);
            slab = (slab / 2 + T - 1) / T * T;     // other parties hold the device's memory: work in smaller passes
            ctx->msm_affine_ws_limit = slab;
        }
    }
    for (uint64_t base = 0; base < entries_up; base += slab) {
        const uint64_t slab_len = std::min<uint64_t>(slab, entries_up - base);
        BaHalf &h0 = half[0];
        h0.base = (uint32_t)base;
        h0.len = (uint32_t)slab_len;
        const int nh = 1;
        k_ba_levels<<<ceil_div_u32(h0.len, 256), 256, 0, st>>>(E_ptr, sorted, h0.base, h0.len, levels, h0.lvl->as<uint8_t>());
        SCZ_LAUNCH_CHECK(ctx);
        for (uint32_t k = 0; k < levels && (h0.len >> (k + 1)); k++) {
            const BaItem it = h0.item(k);
            k_ba_phase1<<<ba_ctas(it), BA_THREADS, 0, st>>>(sp, nseg, sorted, it);
            SCZ_LAUNCH_CHECK(ctx);
            SCZ_TRY(h0.tree.run(ctx, ba_ctas(it) * BA_THREADS));
            k_ba_phase2<<<ba_ctas(it), BA_THREADS, 0, st>>>(sp, nseg, sorted, it);
            SCZ_LAUNCH_CHECK(ctx);
        }
        for (int h = 0; h < nh; h++) {
            const uint32_t chunks = (uint32_t)((half[h].len + T - 1) >> logT);
            k_ba_accumulate<<<ceil_div_u32(chunks, BA_ACC_THREADS), BA_ACC_THREADS, 0, st>>>(
                sp, nseg, E_ptr, logT, sorted, counts, cursor, half[h].lvl->as<uint8_t>(), half[h].base, half[h].len, half[h].L,
                buckets, parts);
            SCZ_LAUNCH_CHECK(ctx);
        }
    }
    return SCZ_OK;
}

}   // namespace scz
