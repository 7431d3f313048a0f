// Batched Pippenger MSM over BLS12-381 G1 (internal interface).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "ctx.h"

namespace scz {

constexpr int MSM_MAX_LEVELS = 7;   // 8^7 = 2^21 buckets per window: window sizes up to 22 bits

// one MSM of a batch ("segment"); lives in device memory, read by every kernel of the pipeline
struct MsmSeg {
    const void *bases;     // packed affine, 96 B per point
    const void *scalars;   // Fr Montgomery, 32 B per scalar
    uint32_t len;
    uint32_t point_base;   // prefix sum of len over the batch
    uint32_t c;            // window bits
    uint32_t W;            // bucket windows: ceil(256 / c), or 1 with a fixed-base table (all digits share one bucket set)
    uint32_t Wd;           // digits per scalar = ceil(256 / c)
    uint32_t pre;          // 1: `bases` is a fixed-base table [digit window][point] (srs.cu)
    uint32_t nb;           // buckets per window = 2^(c-1)
    uint32_t bucket_base;  // first global bucket of this segment (window w starts at bucket_base + w*nb)
    uint32_t window_base;  // first global window of this segment
    // bucket-reduction tree: level k (0-based) has lvl_nodes[k] = max(1, nb >> 3(k+1)) nodes per window, each over
    // (up to) 8 children of the level below (level 0's children are the buckets); the last level has one node
    uint32_t levels;
    uint32_t lvl_nodes[MSM_MAX_LEVELS];
    uint32_t lvl_base[MSM_MAX_LEVELS];   // global node index of (window 0, node 0) at level k
    void *out;             // where this segment's result goes (one Jacobian point); null: slot `k` of d_out_jac
};

#if defined(__CUDACC__)
// the segment whose bucket range holds global bucket id `b`
__device__ __forceinline__ int seg_by_bucket(const MsmSeg *segs, int K, uint32_t b) {
    int lo = 0, hi = K - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&segs[mid].bucket_base) <= b) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
#endif

// The curve-independent front of the pipeline (msm.cu): histogram of the signed digits, exclusive scan, scatter of one
// 8-byte entry (point | sign, bucket) per non-zero digit.  counts (zeroed by the caller) / cursor: one word per bucket.
uint32_t msm_scan_tiles(uint64_t buckets);
int32_t msm_sort_entries(Ctx *ctx, const MsmSeg *d_segs, int K, uint32_t points, uint32_t buckets, uint32_t *counts,
                         uint32_t *cursor, uint32_t *tile_scratch, uint2 *sorted);

// Batched-affine bucket accumulation (msm_affine.cu): replaces the k_msm_accumulate launch of a sequence.  Same
// contract: whole buckets to `buckets` (XYZZ), pieces cut by the 2^logT-entry chunk boundaries to `parts`.
// levels = affine tree levels (1 .. logT).  entries = host-side upper bound of the stream length, *E_ptr the real one.
int32_t msm_accumulate_affine(Ctx *ctx, const MsmSeg *d_segs, int nseg, const uint32_t *E_ptr, uint64_t entries,
                              uint32_t logT, const uint2 *sorted, const uint32_t *counts, const uint32_t *cursor,
                              void *buckets, void *parts, uint32_t levels);

// d_out: batch Jacobian points (144 B each).  Asynchronous on ctx->stream.
// d_outs (optional, host array of `batch` device pointers) sends each result to its own address instead.
// pre_c (optional, host array): per segment the window of a fixed-base table passed as its bases, 0 = plain bases.
int32_t msm_g1_batched(Ctx *ctx, const void *const *d_bases, const void *const *d_scalars, const size_t *lens,
                       size_t batch, void *d_out_jac, void *const *d_outs = nullptr, const uint32_t *pre_c = nullptr);
uint32_t msm_pick_window_pre(size_t len);
uint32_t msm_pick_window(size_t len);

}   // namespace scz
