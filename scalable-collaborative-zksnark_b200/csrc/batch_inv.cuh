// Device-wide batch inversion (Montgomery's trick as a product tree): inverses of n field elements for ONE field
// inversion.  Used by the batched-affine bucket accumulation (msm_affine.cu, Fq: the denominators of a whole tree level)
// and by the point-wise division h = num / den of dhyperplonk.rs:338-339 (poly.cu, Fr).
//   up    thread g owns values g, g + G, g + 2 G, ... (G groups of at most F values): inclusive prefix products inside the
//         group, group total to the next level
//   top   at most INV_TOP values: one thread, one inversion (binary extended Euclid, field.cuh fp_inv_bingcd)
//   down  inverse of value e = (inverse of the group's prefix through e) * (prefix through e - 1), walking backwards
// Values must be non-zero (callers substitute one where no inverse is needed).
#pragma once
#include <memory>
#include <vector>

#include "ctx.h"
#include "field.cuh"

namespace scz {

constexpr int INV_THREADS = 128;
constexpr int INV_F = 32;     // fan-in (values per thread, strided by the number of groups)
constexpr int INV_TOP = 32;   // the tree stops at <= this many values

// Group g owns the values g, g + ngroups, g + 2 ngroups, ... (at most INV_F of them): consecutive lanes touch consecutive
// elements in every step.  (With contiguous groups -- lane stride 32 elements -- every load of a warp touched 32 lines and the
// two big levels of a tree ran at ~150 GB/s.)
template <class P>
__global__ void __launch_bounds__(INV_THREADS) k_inv_tree_up(const void *V, uint32_t n, void *pfx, void *tot, uint32_t ngroups) {
    uint32_t g = blockIdx.x * INV_THREADS + threadIdx.x;
    if (g >= ngroups) return;
    Fp<P> run = Fp<P>::one();
    for (uint32_t e = g; e < n; e += ngroups) {
        run = fp_mul(run, fp_load_rw<P>(V, e));
        fp_store<P>(pfx, e, run);
    }
    fp_store<P>(tot, g, run);
}
template <class P>
__global__ void k_inv_tree_top(const void *V, uint32_t n, void *inv) {
    if (threadIdx.x || blockIdx.x) return;
    Fp<P> pre[INV_TOP];
    Fp<P> run = Fp<P>::one();
    for (uint32_t e = 0; e < n; e++) {
        pre[e] = run;
        run = fp_mul(run, fp_load_rw<P>(V, e));
    }
    Fp<P> I = fp_inv_bingcd(run);
    for (uint32_t e = n; e-- > 0;) {
        fp_store<P>(inv, e, fp_mul(I, pre[e]));
        I = fp_mul(I, fp_load_rw<P>(V, e));
    }
}
// pfx: in = inclusive prefix products of V inside each group, out = the inverse of every value
template <class P>
__global__ void __launch_bounds__(INV_THREADS) k_inv_tree_down(const void *V, uint32_t n, void *pfx, const void *inv_parent,
                                                               uint32_t ngroups) {
    uint32_t g = blockIdx.x * INV_THREADS + threadIdx.x;
    if (g >= ngroups || g >= n) return;
    Fp<P> I = fp_load_rw<P>(inv_parent, g);
    uint32_t e = g + (n - 1 - g) / ngroups * ngroups;   // the group's last element
    while (true) {
        const bool first = e == g;
        Fp<P> inv_e = first ? I : fp_mul(I, fp_load_rw<P>(pfx, e - ngroups));
        if (!first) I = fp_mul(I, fp_load_rw<P>(V, e));
        fp_store<P>(pfx, e, inv_e);
        if (first) break;
        e -= ngroups;
    }
}

// the arrays of a tree over up to `cap` values; values() is written by the caller, inverses() read after run()
template <class P>
struct InvTree {
    std::vector<std::unique_ptr<DevTmp>> V, Pf;
    int32_t alloc(Ctx *ctx, uint64_t cap, bool persistent = false) {
        uint64_t n = cap ? cap : 1;
        while (true) {
            V.emplace_back(new DevTmp(ctx, persistent));
            Pf.emplace_back(new DevTmp(ctx, persistent));
            SCZ_TRY(V.back()->alloc(n * sizeof(Fp<P>)));
            SCZ_TRY(Pf.back()->alloc(n * sizeof(Fp<P>)));
            if (n <= INV_TOP) break;
            n = (n + INV_F - 1) / INV_F;
        }
        return SCZ_OK;
    }
    void *values() const { return V[0]->p; }
    const void *inverses() const { return Pf[0]->p; }
    // inverses()[i] = 1 / values()[i] for i < n0 (n0 <= the capacity given to alloc)
    int32_t run(Ctx *ctx, uint32_t n0) {
        cudaStream_t st = ctx->stream;
        std::vector<uint32_t> n;
        n.push_back(n0);
        size_t lv = 0;
        while (n[lv] > INV_TOP) {
            uint32_t groups = (n[lv] + INV_F - 1) / INV_F;
            k_inv_tree_up<P><<<ceil_div_u32(groups, INV_THREADS), INV_THREADS, 0, st>>>(V[lv]->p, n[lv], Pf[lv]->p, V[lv + 1]->p, groups);
            SCZ_LAUNCH_CHECK(ctx);
            n.push_back(groups);
            lv++;
        }
        k_inv_tree_top<P><<<1, 32, 0, st>>>(V[lv]->p, n[lv], Pf[lv]->p);
        SCZ_LAUNCH_CHECK(ctx);
        while (lv-- > 0) {
            uint32_t groups = n[lv + 1];
            k_inv_tree_down<P><<<ceil_div_u32(groups, INV_THREADS), INV_THREADS, 0, st>>>(V[lv]->p, n[lv], Pf[lv]->p, Pf[lv + 1]->p, groups);
            SCZ_LAUNCH_CHECK(ctx);
        }
        return SCZ_OK;
    }
};

}   // namespace scz
