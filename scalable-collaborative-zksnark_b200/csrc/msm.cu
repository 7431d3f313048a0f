// Batched Pippenger MSM over BLS12-381 G1 for sm_100a.
//
// Replaces `G::msm` (ark-ec 0.4.2 VariableBaseMSM) at the reference's call sites
// dist-primitive/src/dmsm.rs:23 and dpoly_comm.rs:242,274,457.  One launch sequence
// handles a whole batch of independent MSMs ("segments") -- the shape `d_msm`
// receives from `c_open` (22 MSMs of halving length, dpoly_comm.rs:436) -- so
// the ~800 MSMs of a proof cost ~100 sequences instead of ~800.
//
// Pipeline (all on ctx->stream, nothing returns to the host):
//   1 count      one thread per scalar: Montgomery -> integer, signed c-bit digits,
//                histogram of (segment, window, |digit|) buckets               [atomics]
//   2 scan       exclusive prefix sum of the histogram                         [3 small kernels]
//   3 scatter    digits recomputed; (bucket id, point index | sign) written to the bucket's
//                slot: a counting sort whose within-bucket order is irrelevant because
//                group addition commutes (the result is bit-exact regardless)
//   4 accumulate the sorted entry stream is cut into equal chunks of T entries, one thread per chunk: perfectly
//                balanced whatever the digit distribution (the all-ones scalars of dmsm.rs:103 included).  Big streams
//                with long bucket runs take the batched-affine path (msm_affine.cu: affine additions with one shared
//                inversion per tree level inside the aligned single-bucket blocks of the stream, XYZZ mixed additions
//                for what is left); small ones the plain XYZZ kernel below (XYZZ += affine per entry, 96 B gathered
//                with 128-bit loads, next point prefetched during the current add).  Either way runs that are whole
//                buckets go straight to the bucket array, the first / last run of a chunk may be a piece of a bucket
//                shared with the neighbours
//   5 fix-up     pieces of buckets that straddle chunk boundaries are summed
//   6 tree       bucket reduction sum_j (j + 1) B_j per window as a tree of fan-in 8: a node keeps S = sum B and
//                T = sum (j - base) B; two general additions per bucket at level 0, three per node above
//   7 finish     per segment: one Horner chain over the window sums (groups of 4 cooperating lanes), XYZZ -> Jacobian;
//                with a fixed-base table there is ONE window sum and no chain
// The XYZZ kernel is bound by the integer multiply pipe (a mixed add is 2 736 IMAD.WIDE for ~104 B of HBM traffic), the
// affine levels by the same pipe (upper levels) and by the DRAM rate of random gathers (level 0), see DESIGN.md 3.1.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "deferred.h"
#include "g1_coop.cuh"
#include "msm.h"
#include "msm_digits.cuh"

namespace scz {

constexpr int CNT_THREADS = 256;
constexpr int ACC_THREADS = 128;
constexpr int FIX_THREADS = 128;
constexpr int CHK_THREADS = 64;
constexpr int PLN_THREADS = 256;
constexpr int FIN_THREADS = 256;       // >= max windows of a segment (c = 1 -> 256)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;          // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int seg_by_point(const MsmSeg *segs, int K, uint32_t g) {
    int lo = 0, hi = K - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&segs[mid].point_base) <= g) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ int seg_by_window(const MsmSeg *segs, int K, uint32_t w) {
    int lo = 0, hi = K - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&segs[mid].window_base) <= w) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ int seg_by_level(const MsmSeg *segs, int K, int lvl, uint32_t t) {
    int lo = 0, hi = K - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&segs[mid].lvl_base[lvl]) <= t) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// Out-of-line group law for the latency-bound tail kernels: keeps their code inside the
// instruction cache (an inlined general add is ~5k instructions) and the build fast.
__device__ __noinline__ Fq fq_mul_nl(const Fq &a, const Fq &b) { return fp_mul(a, b); }
__device__ __noinline__ Fq fq_sqr_nl(const Fq &a) { return fp_sqr(a); }
__device__ __noinline__ Fq fq_dot2_sub_nl(const Fq &a, const Fq &b, const Fq &c, const Fq &d) { return fp_dot2_sub(a, b, c, d); }
struct MulCall {
    __device__ __forceinline__ static Fq mul(const Fq &a, const Fq &b) { return fq_mul_nl(a, b); }
    __device__ __forceinline__ static Fq sqr(const Fq &a) { return fq_sqr_nl(a); }
    __device__ __forceinline__ static Fq dot2_sub(const Fq &a, const Fq &b, const Fq &c, const Fq &d) {
        return fq_dot2_sub_nl(a, b, c, d);
    }
};
// (measured, not kept: ONE out-of-line addition with its 14 products inlined, ~85 KB of straight-line code, instead of
// out-of-line products: bucket tree 14.5 vs 13.9 ms per 2^20 proof)
__device__ __noinline__ void g1x_add_nl(G1X &r, const G1X &a, const G1X &b) { r = g1x_add<MulCall>(a, b); }
__device__ __noinline__ void g1x_double_nl(G1X &r, const G1X &a) { r = g1x_double<MulCall>(a); }

// ---- 1 + 3: recode, then count (SCATTER = false) or place (SCATTER = true)
template <bool SCATTER>
__global__ void __launch_bounds__(CNT_THREADS) k_msm_recode(const MsmSeg *segs, int K, uint32_t total_points,
                                                             uint32_t *counts, uint32_t *cursor, uint2 *sorted) {
    uint32_t g = blockIdx.x * CNT_THREADS + threadIdx.x;
    if (g >= total_points) return;
    int s = K == 1 ? 0 : seg_by_point(segs, K, g);
    const MsmSeg sg = segs[s];
    uint32_t i = g - sg.point_base;
    Fr k = fp_to_canon(fp_load<FrP>(sg.scalars, i));
    uint32_t carry = 0;
    for (uint32_t w = 0; w < sg.Wd; w++) {
        int32_t d = msm_signed_digit(k.l, sg.c, w, carry);
        if (d == 0) continue;
        uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        // with a fixed-base table every window lands in the same bucket set and refers to the multiple 2^(c w) P_i
        uint32_t b = sg.bucket_base + (sg.pre ? 0 : w * sg.nb) + (mag - 1);
        uint32_t ref = sg.pre ? w * sg.len + i : i;
        if (!SCATTER) {
            atomicAdd(&counts[b], 1u);
        } else {
            uint32_t pos = atomicAdd(&cursor[b], 1u);
            sorted[pos] = make_uint2(ref | (d < 0 ? 0x80000000u : 0u), b);   // one 8-byte store: (point | sign, bucket)
        }
    }
}

// ---- 2: exclusive scan of counts -> cursor (three passes)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t *in, uint32_t *out, uint32_t *tile_sums,
                                                             uint32_t n) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        v[j] = base + j < n ? in[base + j] : 0;
        sum += v[j];
    }
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0, wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - ws;   // exclusive
        if (lane == SCAN_THREADS / 32 - 1) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[wid] + incl - sum;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
}
__global__ void __launch_bounds__(1024) k_scan_tile_sums(uint32_t *tile_sums, uint32_t tiles) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t start = 0; start < tiles; start += 1024) {
        uint32_t idx = start + threadIdx.x;
        uint32_t v = idx < tiles ? tile_sums[idx] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - ws;
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + warp_sums[wid] + incl - v;
        if (idx < tiles) tile_sums[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t *out, const uint32_t *tile_sums, uint32_t n) {
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++)
        if (base + j < n) out[base + j] += add;
}

// ---- 4: bucket accumulation over equal chunks of the sorted entry stream
// After the scatter pass cursor[b] is the END of bucket b; it starts at cursor[b] - counts[b].
// Chunk t covers entries [t*T, (t+1)*T).  A bucket that lies inside one chunk is written to
// buckets[b]; a bucket cut by chunk boundaries leaves a TAIL piece in the chunk where it starts
// (parts[2t+1]) and HEAD pieces in the following chunks (parts[2u]); k_msm_fixup* sums them.
#ifndef ACC_MIN_BLOCKS
#define ACC_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(ACC_THREADS, ACC_MIN_BLOCKS) k_msm_accumulate(const MsmSeg *segs, int K, const uint32_t *E_ptr,
                                                                 uint32_t logT, const uint2 *sorted,
                                                                 const uint32_t *counts,
                                                                 const uint32_t *cursor, void *buckets, void *parts) {
    uint32_t t = blockIdx.x * ACC_THREADS + threadIdx.x;
    const uint32_t E = __ldg(E_ptr);   // entries actually in the stream (zero digits are dropped)
    uint64_t lo64 = (uint64_t)t << logT;
    if (lo64 >= E) return;
    uint32_t lo = (uint32_t)lo64;
    uint32_t hi = lo64 + (1u << logT) < E ? lo + (1u << logT) : E;
    const uint2 e_first = __ldg(sorted + lo);
    uint32_t k_first = e_first.y, k_last = __ldg(sorted + hi - 1).y;
    bool head_piece = cursor[k_first] - counts[k_first] < lo;   // the first bucket began in an earlier chunk
    bool tail_piece = cursor[k_last] > hi;                      // the last bucket goes on in a later chunk

    // segment of the current run (bases pointer); re-resolved when the bucket id leaves its range
    uint32_t seg_hi = 0;
    const void *bases = nullptr;
    uint32_t cur = k_first;
    bool first_run = true;
    G1X acc = G1X::inf();
    {
        int s = K == 1 ? 0 : seg_by_bucket(segs, K, cur);
        seg_hi = segs[s].bucket_base + segs[s].W * segs[s].nb;
        bases = segs[s].bases;
    }
    // Software pipeline, two entries deep: while entry j is added, the POINT of entry j+1 is gathered (its index
    // arrived one iteration ago, so the gather issues at once) and the ENTRY j+2 is fetched.  Nothing on the
    // index -> address -> point chain is ever waited for at the top of an iteration.
    uint32_t ent = e_first.x;
    G1Affine p = g1a_load_stream(bases, ent & 0x7fffffffu);
    uint2 ne = lo + 1 < hi ? __ldg(sorted + lo + 1) : make_uint2(0, cur);
    for (uint32_t j = lo; j < hi; j++) {
        const bool more = j + 1 < hi;
        const uint32_t nkey = more ? ne.y : cur, nent = ne.x;
        G1Affine np;
        uint2 n2 = make_uint2(0, 0);
        if (j + 2 < hi) n2 = __ldg(sorted + j + 2);
        if (more) {
            if (nkey >= seg_hi) {
                int s = seg_by_bucket(segs, K, nkey);
                seg_hi = segs[s].bucket_base + segs[s].W * segs[s].nb;
                bases = segs[s].bases;
            }
            np = g1a_load_stream(bases, nent & 0x7fffffffu);
        }
        g1x_add_affine(acc, p, (ent >> 31) != 0);   // fully inlined: out-of-line products cost 60 % here (measured)
        if (!more || nkey != cur) {   // the run of bucket `cur` ends here
            bool last_run = !more;
            if (first_run && head_piece) g1x_store(parts, 2 * (size_t)t, acc);
            else if (last_run && tail_piece) g1x_store(parts, 2 * (size_t)t + 1, acc);
            else g1x_store(buckets, cur, acc);
            first_run = false;
            acc = G1X::inf();
            cur = nkey;
        }
        if (more) {
            p = np;
            ent = nent;
            ne = n2;
        }
    }
}

// ---- 5: buckets cut by chunk boundaries.  Bucket b = entries [start, end) touches chunks t0 = start/T ..
// t1 = (end-1)/T; when t1 > t0 its value is parts[2*t0+1] + sum_{u in (t0, t1]} parts[2u].  One thread per CHUNK:
// chunk t owns the bucket that starts in it and runs past its end (nearly every chunk has one, so the kernel is
// dense).  Long chains (the top window's few buckets, degenerate scalar distributions such as dmsm.rs:103) go
// to a list served by whole CTAs.
constexpr uint32_t FIX_SERIAL_MAX = 6;
__global__ void __launch_bounds__(FIX_THREADS) k_msm_fixup(const uint32_t *E_ptr, uint32_t logT, const uint2 *sorted,
                                                            const uint32_t *counts, const uint32_t *cursor,
                                                            const void *parts, void *buckets, uint32_t *heavy_list,
                                                            uint32_t *heavy_count) {
    uint32_t t = blockIdx.x * FIX_THREADS + threadIdx.x;
    const uint32_t E = __ldg(E_ptr);
    uint64_t lo64 = (uint64_t)t << logT;
    if (lo64 >= E) return;
    uint32_t lo = (uint32_t)lo64;
    uint32_t hi = lo64 + (1u << logT) < E ? lo + (1u << logT) : E;
    uint32_t b = __ldg(sorted + hi - 1).y;
    uint32_t end = cursor[b];
    if (end <= hi) return;                      // the chunk's last bucket ends inside it
    uint32_t start = end - counts[b];
    if (start < lo) return;                     // began in an earlier chunk: that chunk owns it
    uint32_t t1 = (end - 1) >> logT;
    if (t1 - t > FIX_SERIAL_MAX) {
        heavy_list[atomicAdd(heavy_count, 1u)] = b;
        return;
    }
    G1X acc = g1x_load(parts, 2 * (size_t)t + 1);
    for (uint32_t u = t + 1; u <= t1; u++) {
        G1X h = g1x_load(parts, 2 * (size_t)u);
        g1x_add_nl(acc, acc, h);
    }
    g1x_store(buckets, b, acc);
}

// ---- 6: bucket reduction, sum_j (j+1) B_j per window, as a tree of fan-in 8.  A node over the bucket range
//   [base, base + span) keeps S = sum B_j and T = sum (j - base) B_j.  A parent over children c_0 .. c_7 (each of
//   span s): S = sum S_i (running sum, high child first), R = sum_i i*S_i (sum of the running sums), T = sum_i T_i + s*R
//   (s = 8^level: 3*level doublings).  Two general additions per bucket at level 0, three per node above:
//   ~2.4 additions per bucket in total, all in dense thread-per-node kernels.  The root writes S + T, the window sum.
struct MsmNode {
    G1X S, T;
};
template <bool FIRST>
__global__ void __launch_bounds__(CHK_THREADS) k_msm_tree(const MsmSeg *segs, int K, int lvl, uint32_t first, uint32_t count,
                                                           const uint32_t *counts, const void *buckets, MsmNode *nodes,
                                                           void *wsum) {
    uint32_t t = blockIdx.x * CHK_THREADS + threadIdx.x;
    if (t >= count) return;
    t += first;                                        // global node index
    int s = K == 1 ? 0 : seg_by_level(segs, K, lvl, t);
    const MsmSeg sg = segs[s];
    uint32_t per_w = sg.lvl_nodes[lvl];
    uint32_t rel = t - sg.lvl_base[lvl];
    uint32_t w = rel / per_w, i = rel - w * per_w;
    uint32_t below = FIRST ? sg.nb : sg.lvl_nodes[lvl - 1];
    uint32_t nchild = below / per_w;                  // 8, or fewer at the root / for tiny windows
    G1X S = G1X::inf(), R = G1X::inf(), Tsum = G1X::inf();
    if (FIRST) {
        uint32_t b0 = sg.bucket_base + w * sg.nb + i * nchild;
        for (uint32_t j = nchild; j-- > 0;) {
            if (counts[b0 + j]) {
                G1X b = g1x_load(buckets, b0 + j);
                g1x_add_nl(S, S, b);
            }
            if (j) g1x_add_nl(R, R, S);
        }
    } else {
        const MsmNode *ch = nodes + sg.lvl_base[lvl - 1] + (size_t)w * below + (size_t)i * nchild;
        for (uint32_t j = nchild; j-- > 0;) {
            G1X cs = g1x_load(&ch[j].S, 0), ct = g1x_load(&ch[j].T, 0);
            g1x_add_nl(S, S, cs);
            g1x_add_nl(Tsum, Tsum, ct);
            if (j) g1x_add_nl(R, R, S);
        }
        for (int d = 0; d < 3 * lvl; d++) g1x_double_nl(R, R);
        g1x_add_nl(R, R, Tsum);
    }
    if (lvl + 1 == (int)sg.levels) {                   // root: window sum = sum_j (j + 1) B_j = T + S
        g1x_add_nl(R, R, S);
        g1j_store(wsum, sg.window_base + w, g1x_to_jac(R));
    } else {
        g1x_store(&nodes[t].S, 0, S);
        g1x_store(&nodes[t].T, 0, R);
    }
}

// tree-sum of one XYZZ value per thread through shared memory; result in thread 0
template <int T>
__device__ __forceinline__ G1X block_sum_g1x(G1X v, G1X *sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int stride = T / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
            G1X o = sh[threadIdx.x + stride];
            if (!o.is_inf()) {
                g1x_add_nl(v, v, o);
                sh[threadIdx.x] = v;
            }
        }
        __syncthreads();
    }
    return v;
}

// Long chains: one WARP per bucket while the chain is short enough for 32 lanes (the few hundred low buckets that
// collect a fixed-base table's narrow top window), the whole CTA for the really long ones (degenerate scalars).
constexpr uint32_t FIX_WARP_MAX = 256;   // parts; beyond this the CTA sums the chain together
__global__ void __launch_bounds__(PLN_THREADS) k_msm_fixup_heavy(uint32_t logT, const uint32_t *counts,
                                                                  const uint32_t *cursor, const void *parts,
                                                                  void *buckets, const uint32_t *heavy_list,
                                                                  const uint32_t *heavy_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    G1X *sh = reinterpret_cast<G1X *>(smem_raw);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = PLN_THREADS / 32;
    const uint32_t total = *heavy_count;
    // pass 1: warp per bucket
    for (uint32_t h = blockIdx.x * nw + wid; h < total; h += gridDim.x * nw) {
        uint32_t b = heavy_list[h];
        uint32_t end = cursor[b], start = end - counts[b];
        uint32_t t0 = start >> logT, t1 = (end - 1) >> logT;
        if (t1 - t0 > FIX_WARP_MAX) continue;
        G1X acc = G1X::inf();
        if (lane == 0) acc = g1x_load(parts, 2 * (size_t)t0 + 1);
        for (uint32_t u = t0 + 1 + lane; u <= t1; u += 32) {
            G1X v = g1x_load(parts, 2 * (size_t)u);
            g1x_add_nl(acc, acc, v);
        }
        G1X *w = sh + wid * 32;
        w[lane] = acc;
        __syncwarp();
        for (int stride = 16; stride > 0; stride >>= 1) {
            if ((int)lane < stride) {
                G1X o = w[lane + stride];
                if (!o.is_inf()) {
                    g1x_add_nl(acc, acc, o);
                    w[lane] = acc;
                }
            }
            __syncwarp();
        }
        if (lane == 0) g1x_store(buckets, b, acc);
        __syncwarp();
    }
    __syncthreads();
    // pass 2: CTA per bucket for the very long chains
    for (uint32_t h = blockIdx.x; h < total; h += gridDim.x) {
        uint32_t b = heavy_list[h];
        uint32_t end = cursor[b], start = end - counts[b];
        uint32_t t0 = start >> logT, t1 = (end - 1) >> logT;
        if (t1 - t0 <= FIX_WARP_MAX) continue;   // uniform across the CTA
        G1X acc = G1X::inf();
        if (threadIdx.x == 0) acc = g1x_load(parts, 2 * (size_t)t0 + 1);
        for (uint32_t u = t0 + 1 + threadIdx.x; u <= t1; u += PLN_THREADS) {
            G1X v = g1x_load(parts, 2 * (size_t)u);
            g1x_add_nl(acc, acc, v);
        }
        acc = block_sum_g1x<PLN_THREADS>(acc, sh);
        if (threadIdx.x == 0) g1x_store(buckets, b, acc);
        __syncthreads();
    }
}

// ---- 7: one CTA per segment: Horner over the window sums (255 doublings: the longest chain of the pipeline),
//         run by groups of 4 cooperating lanes (g1_coop.cuh); warp 0's 8 groups run the same chain redundantly.
__global__ void __launch_bounds__(FIN_THREADS) k_msm_finish(const MsmSeg *segs, const void *wsum, void *out_jac) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    G1Jac *sh = reinterpret_cast<G1Jac *>(smem_raw);   // [W]
    const MsmSeg sg = segs[blockIdx.x];
    const Coop g;
    if (sg.len)
        for (uint32_t w = threadIdx.x; w < sg.W; w += FIN_THREADS) sh[w] = g1j_load(wsum, sg.window_base + w);
    __syncthreads();
    if (threadIdx.x < 32) {
        G1Jac acc = g1j_inf();
        if (sg.len) {
            for (uint32_t ww = sg.W; ww-- > 0;) {
                if (!acc.z.is_zero())
                    for (uint32_t i = 0; i < sg.c; i++) coop_double(g, acc);
                G1Jac v = sh[ww];
                coop_add(g, acc, v);
            }
        }
        if (threadIdx.x == 0) {
            if (sg.out) g1j_store(sg.out, 0, acc);
            else g1j_store(out_jac, blockIdx.x, acc);
        }
    }
}

// ------------------------------------------------------------------ host side
uint32_t msm_pick_window(size_t len) {
    if (len == 0) return 1;
    // cost in mixed-add equivalents: W * (len + 3 * nb): a bucket costs two general additions in
    // the level-1 reduction
    uint32_t best = 1;
    double best_cost = 1e300;
    for (uint32_t c = 1; c <= 16; c++) {
        double W = (double)msm_num_windows(c), nb = (double)(1u << (c - 1));
        double cost = W * ((double)len + 3.0 * nb);
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

// Steps 1-3 for a batch described by `d_segs` (device): histogram, exclusive scan, scatter.  Curve-independent (only the
// scalars are read): the G2 path (msm_g2.cu) sorts with it too.  counts was zeroed by the caller; afterwards cursor[b] is the
// END of bucket b in `sorted`, counts[b] its length.  tiles: ceil(buckets / SCAN_TILE) + 1 words of scratch.
uint32_t msm_scan_tiles(uint64_t buckets) { return ceil_div_u32(buckets, SCAN_TILE); }
int32_t msm_sort_entries(Ctx *ctx, const MsmSeg *d_segs, int K, uint32_t points, uint32_t buckets, uint32_t *counts,
                         uint32_t *cursor, uint32_t *tile_scratch, uint2 *sorted) {
    ProfScope ps(ctx, SCZ_K_MSM_SORT);
    cudaStream_t st = ctx->stream;
    const uint32_t tiles = msm_scan_tiles(buckets);
    if (points) {
        k_msm_recode<false><<<ceil_div_u32(points, CNT_THREADS), CNT_THREADS, 0, st>>>(d_segs, K, points, counts, nullptr, nullptr);
        SCZ_LAUNCH_CHECK(ctx);
    }
    k_scan_tiles<<<tiles, SCAN_THREADS, 0, st>>>(counts, cursor, tile_scratch, buckets);
    SCZ_LAUNCH_CHECK(ctx);
    k_scan_tile_sums<<<1, 1024, 0, st>>>(tile_scratch, tiles);
    SCZ_LAUNCH_CHECK(ctx);
    k_scan_add<<<tiles, SCAN_THREADS, 0, st>>>(cursor, tile_scratch, buckets);
    SCZ_LAUNCH_CHECK(ctx);
    if (points) {
        // (measured, not kept: a persisting L2 access-policy window over the cursor array during this pass, whose slot
        // reservations hit it at random while GBs of 8-byte stores stream through L2 -- sort 12.1 / 17.8 ms per 2^20 proof
        // with a 32 / 64 MiB set-aside against 10.5 ms without)
        k_msm_recode<true><<<ceil_div_u32(points, CNT_THREADS), CNT_THREADS, 0, st>>>(d_segs, K, points, nullptr, cursor, sorted);
        SCZ_LAUNCH_CHECK(ctx);
    }
    return SCZ_OK;
}

int32_t msm_g1_batched(Ctx *ctx, const void *const *d_bases, const void *const *d_scalars, const size_t *lens,
                       size_t batch, void *d_out, void *const *d_outs, const uint32_t *pre_c) {
    if (batch == 0) return SCZ_OK;
    if (batch > (1u << 20)) return ctx->fail(SCZ_ERR_BAD_ARG, "msm: batch too large");
    std::vector<MsmSeg> segs(batch);
    uint64_t points = 0, buckets = 0, windows = 0, entries = 0, lvl_count[MSM_MAX_LEVELS] = {0};
    uint32_t max_levels = 1;
    for (size_t k = 0; k < batch; k++) {
        if (lens[k] >= (1ull << 31)) return ctx->fail(SCZ_ERR_BAD_ARG, "msm: segment %zu too long", k);
        if (lens[k] && (!d_bases[k] || !d_scalars[k])) return ctx->fail(SCZ_ERR_BAD_ARG, "msm: null segment %zu", k);
        MsmSeg &s = segs[k];
        s.bases = d_bases[k];
        s.scalars = d_scalars[k];
        s.len = (uint32_t)lens[k];
        s.point_base = (uint32_t)points;
        s.pre = pre_c && pre_c[k] ? 1 : 0;
        s.c = s.pre ? pre_c[k] : (ctx->msm_window_override ? ctx->msm_window_override : msm_pick_window(lens[k]));
        s.Wd = msm_num_windows(s.c);
        s.W = s.pre ? 1 : s.Wd;
        if ((uint64_t)s.len * s.Wd >= (1ull << 31)) return ctx->fail(SCZ_ERR_BAD_ARG, "msm: segment %zu too long", k);
        s.nb = 1u << (s.c - 1);
        s.bucket_base = (uint32_t)buckets;
        s.window_base = (uint32_t)windows;
        s.levels = 0;
        for (uint32_t below = s.nb; s.levels == 0 || below > 1;) {
            if (s.levels == MSM_MAX_LEVELS) return ctx->fail(SCZ_ERR_BAD_ARG, "msm: window of %u bits is too wide", s.c);
            uint32_t here = std::max<uint32_t>(1, below >> 3);
            s.lvl_nodes[s.levels] = here;
            lvl_count[s.levels] += (uint64_t)s.W * here;
            s.levels++;
            below = here;
        }
        for (uint32_t k2 = s.levels; k2 < MSM_MAX_LEVELS; k2++) s.lvl_nodes[k2] = 0;
        max_levels = std::max(max_levels, s.levels);
        s.out = d_outs ? d_outs[k] : nullptr;
        points += s.len;
        buckets += (uint64_t)s.W * s.nb;
        windows += s.W;
        entries += (uint64_t)s.len * s.Wd;
    }
    // node index space: level 0 of all segments, then level 1 of all segments, ...
    uint64_t lvl_first[MSM_MAX_LEVELS + 1] = {0};
    for (int k2 = 0; k2 < MSM_MAX_LEVELS; k2++) lvl_first[k2 + 1] = lvl_first[k2] + lvl_count[k2];
    {
        uint64_t run[MSM_MAX_LEVELS];
        for (int k2 = 0; k2 < MSM_MAX_LEVELS; k2++) run[k2] = lvl_first[k2];
        for (size_t k = 0; k < batch; k++)
            for (int k2 = 0; k2 < MSM_MAX_LEVELS; k2++) {
                segs[k].lvl_base[k2] = (uint32_t)run[k2];
                run[k2] += (uint64_t)segs[k].W * segs[k].lvl_nodes[k2];
            }
    }
    const uint64_t nodes_total = lvl_first[MSM_MAX_LEVELS];
    if (points >= (1ull << 31) || buckets >= (1ull << 31) || entries >= (1ull << 32) || nodes_total >= (1ull << 32))
        return ctx->fail(SCZ_ERR_BAD_ARG, "msm: batch too large (%llu points, %llu buckets)", (unsigned long long)points,
                         (unsigned long long)buckets);
    ctx->msm_bucket_adds = entries;
    ctx->msm_buckets = buckets;
    ctx->msm_windows = windows;
    ctx->msm_cum_adds += entries, ctx->msm_cum_pairs += points, ctx->msm_cum_sequences++, ctx->msm_cum_segments += batch;

    // entries per accumulate thread: 256 for a whole proof's batch (fewer chunk boundaries = fewer pieces to fix up:
    // 8.3 -> 1.7 ms at 375 M entries), fewer when that would leave less than ~6 waves of CTAs
    uint32_t logT = 8;
    while (logT > 3 && (entries >> logT) < (uint64_t)ctx->sm_count * 2 * ACC_THREADS * 6) logT--;
    uint32_t nchunks = (uint32_t)((entries + (1u << logT) - 1) >> logT);   // upper bound: zero digits shrink the stream

    cudaStream_t st = ctx->stream;
    uint32_t tiles = ceil_div_u32(buckets, SCAN_TILE);
    DevTmp d_segs(ctx), d_counts(ctx), d_cursor(ctx), d_tiles(ctx), d_sorted(ctx), d_buckets(ctx),
        d_parts(ctx), d_heavy(ctx), d_nodes(ctx), d_wsum(ctx);
    SCZ_TRY(d_segs.alloc(batch * sizeof(MsmSeg)));
    SCZ_TRY(d_counts.alloc(buckets * 4));
    SCZ_TRY(d_cursor.alloc(buckets * 4));
    SCZ_TRY(d_tiles.alloc((size_t)tiles * 4 + 4));
    SCZ_TRY(d_sorted.alloc((entries ? entries : 1) * sizeof(uint2)));
    SCZ_TRY(d_buckets.alloc(buckets * sizeof(G1X)));
    SCZ_TRY(d_parts.alloc(((size_t)nchunks * 2 + 2) * sizeof(G1X)));
    SCZ_TRY(d_heavy.alloc(((size_t)nchunks + 1) * 4));   // [0] = count, [1..] = list (at most one bucket per chunk)
    SCZ_TRY(d_nodes.alloc(nodes_total * sizeof(MsmNode)));
    SCZ_TRY(d_wsum.alloc(windows * sizeof(G1Jac)));
    SCZ_TRY(ctx->h2d_staged(d_segs.p, segs.data(), batch * sizeof(MsmSeg)));   // asynchronous: the host runs ahead
    SCZ_CUDA(ctx, cudaMemsetAsync(d_counts.p, 0, buckets * 4, st));
    SCZ_CUDA(ctx, cudaMemsetAsync(d_heavy.p, 0, 4, st));
    const MsmSeg *sp = d_segs.as<MsmSeg>();
    int K = (int)batch;
    uint32_t *counts = d_counts.as<uint32_t>(), *cursor = d_cursor.as<uint32_t>();
    uint2 *sorted = d_sorted.as<uint2>();
    if (!ctx->attr_msm) {   // per device, so per ctx
        SCZ_CUDA(ctx, cudaFuncSetAttribute(k_msm_fixup_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(PLN_THREADS * sizeof(G1X))));
        SCZ_CUDA(ctx, cudaFuncSetAttribute(k_msm_finish, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(FIN_THREADS * sizeof(G1Jac))));
        ctx->attr_msm = true;
    }
    SCZ_TRY(msm_sort_entries(ctx, sp, K, (uint32_t)points, (uint32_t)buckets, counts, cursor, d_tiles.as<uint32_t>(), sorted));
    if (points) {
        {
            // after the scatter cursor[last bucket] = entries really in the stream (the host only knows the
            // bound `entries`: zero digits are dropped); chunks past that end return at once
            ProfScope ps(ctx, SCZ_K_MSM_ACCUMULATE);
            // long bucket runs in a big stream: affine additions with a shared inversion inside the aligned pure blocks
            // of the stream, XYZZ for what is left (msm_affine.cu); otherwise XYZZ for every entry
            uint32_t aff_levels = 0;
            if (ctx->msm_affine_mode != 2) {
                double avg_run = (double)entries / (double)std::max<uint64_t>(1, buckets);
                uint32_t want = ctx->msm_affine_levels;
                if (!want)
                    while (want < 6 && (double)(4u << want) <= avg_run && (entries >> (want + 1)) >= (1ull << 21)) want++;
                want = std::min(want, logT);
                if (ctx->msm_affine_mode == 1) aff_levels = std::max(1u, want);
                else if (want >= 2 && entries >= (1ull << 24)) aff_levels = want;
            }
            if (aff_levels) {
                ctx->msm_cum_affine_sequences++;
                SCZ_TRY(msm_accumulate_affine(ctx, sp, K, cursor + (buckets - 1), entries, logT, sorted, counts, cursor,
                                              d_buckets.p, d_parts.p, aff_levels));
            } else {
                k_msm_accumulate<<<ceil_div_u32(nchunks, ACC_THREADS), ACC_THREADS, 0, st>>>(
                    sp, K, cursor + (buckets - 1), logT, sorted, counts, cursor, d_buckets.p, d_parts.p);
                SCZ_LAUNCH_CHECK(ctx);
            }
        }
        ProfScope ps(ctx, SCZ_K_MSM_FIXUP);
        uint32_t *heavy = d_heavy.as<uint32_t>();
        k_msm_fixup<<<ceil_div_u32(nchunks, FIX_THREADS), FIX_THREADS, 0, st>>>(cursor + (buckets - 1), logT, sorted, counts,
                                                                                cursor, d_parts.p, d_buckets.p, heavy + 1,
                                                                                heavy);
        SCZ_LAUNCH_CHECK(ctx);
        k_msm_fixup_heavy<<<ctx->sm_count * 2, PLN_THREADS, PLN_THREADS * sizeof(G1X), st>>>(
            logT, counts, cursor, d_parts.p, d_buckets.p, heavy + 1, heavy);
        SCZ_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, SCZ_K_MSM_REDUCE);
        for (uint32_t lv = 0; lv < max_levels; lv++) {
            uint32_t cnt = (uint32_t)lvl_count[lv];
            if (!cnt) continue;
            if (lv == 0)
                k_msm_tree<true><<<ceil_div_u32(cnt, CHK_THREADS), CHK_THREADS, 0, st>>>(
                    sp, K, 0, 0, cnt, counts, d_buckets.p, d_nodes.as<MsmNode>(), d_wsum.p);
            else
                k_msm_tree<false><<<ceil_div_u32(cnt, CHK_THREADS), CHK_THREADS, 0, st>>>(
                    sp, K, (int)lv, (uint32_t)lvl_first[lv], cnt, counts, d_buckets.p, d_nodes.as<MsmNode>(), d_wsum.p);
            SCZ_LAUNCH_CHECK(ctx);
        }
    }
    {
        ProfScope ps(ctx, SCZ_K_MSM_FINISH);
        k_msm_finish<<<(uint32_t)batch, FIN_THREADS, FIN_THREADS * sizeof(G1Jac), st>>>(sp, d_wsum.p, d_out);
        SCZ_LAUNCH_CHECK(ctx);
    }
    return SCZ_OK;
}

// ------------------------------------------------------------------ deferred execution (deferred.h)
// One launch sequence is bounded by 32-bit entry / point indices; stay well inside and flush early otherwise.
constexpr uint64_t DEFER_MAX_ENTRIES = 3ull << 30, DEFER_MAX_POINTS = 1ull << 30, DEFER_MAX_SEGS = 1u << 16;

int32_t Deferred::add_msm(const void *b, const void *s, size_t len, void *out, uint32_t pre_c) {
    if (!out || (len && (!b || !s))) return ctx->fail(SCZ_ERR_BAD_ARG, "msm: null argument");
    uint32_t c = pre_c ? pre_c : (ctx->msm_window_override ? ctx->msm_window_override : msm_pick_window(len));
    uint64_t e = (uint64_t)len * msm_num_windows(c);
    if (!lens.empty() && (entries + e > DEFER_MAX_ENTRIES || points + len > DEFER_MAX_POINTS || lens.size() >= DEFER_MAX_SEGS))
        SCZ_TRY(flush_msm());
    bases.push_back(b);
    scalars.push_back(s);
    lens.push_back(len);
    outs.push_back(out);
    pre.push_back(pre_c);
    entries += e;
    points += len;
    return SCZ_OK;
}
// With SCZ_MSM_STREAM=1 the launch sequences run on a second, lowest-priority stream of the ctx, fenced by events on
// both sides (the host gives the ctx's main stream a higher priority).  Two things follow:
//  * flush_early(): the schedule (hyperplonk.cu) starts the sequence of everything queued so far while its own
//    protocol phase is still running -- the latency-bound short kernels of the remaining sumchecks / folds are
//    dispatched ahead of the bucket kernel's queued CTAs and hide under it; run() joins before anything reads a
//    result.  Inputs of queued MSMs are immutable until run() anyway (that is what deferring them relies on).
//  * when several provers share a GPU (proofs in flight on different ctxs) one prover's protocol kernels overlap
//    another's bucket kernel instead of waiting for its whole grid to drain.
static bool msm_side_stream() {
    static const bool side = [] { const char *e = getenv("SCZ_MSM_STREAM"); return e && e[0] == '1'; }();
    return side;
}
static int32_t msm_stream_ready(Ctx *ctx) {
    if (ctx->msm_stream) return SCZ_OK;
    int least = 0, greatest = 0;
    SCZ_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
    SCZ_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->msm_stream, cudaStreamNonBlocking, least));
    SCZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->msm_fork, cudaEventDisableTiming));
    SCZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->msm_join, cudaEventDisableTiming));
    return SCZ_OK;
}
int32_t Deferred::join_early() {
    if (!early_pending) return SCZ_OK;
    early_pending = false;
    SCZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->msm_join, 0));
    return SCZ_OK;
}
Deferred::~Deferred() {
    if (early_pending) cudaStreamWaitEvent(ctx->stream, ctx->msm_join, 0);   // error path: `keep` is freed on the main stream
}
int32_t Deferred::flush_msm_on_side(bool wait) {
    cudaStream_t main_stream = ctx->stream;
    SCZ_TRY(msm_stream_ready(ctx));
    SCZ_CUDA(ctx, cudaEventRecord(ctx->msm_fork, main_stream));
    SCZ_CUDA(ctx, cudaStreamWaitEvent(ctx->msm_stream, ctx->msm_fork, 0));
    ctx->stream = ctx->msm_stream;
    int32_t rc = msm_g1_batched(ctx, bases.data(), scalars.data(), lens.data(), lens.size(), nullptr, outs.data(), pre.data());
    ctx->stream = main_stream;
    bases.clear(), scalars.clear(), lens.clear(), outs.clear(), pre.clear();
    entries = points = 0;
    SCZ_CUDA(ctx, cudaEventRecord(ctx->msm_join, ctx->msm_stream));
    if (wait) SCZ_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->msm_join, 0));
    else early_pending = true;
    return rc;
}
int32_t Deferred::flush_early() {
    if (!msm_side_stream() || lens.empty()) return SCZ_OK;
    SCZ_TRY(join_early());
    early_after = after.size();      // every continuation registered so far has its MSMs in this sequence
    early_after2 = after2.size();
    return flush_msm_on_side(false);
}
int32_t Deferred::flush_msm() {
    SCZ_TRY(join_early());
    if (lens.empty()) return SCZ_OK;
    if (msm_side_stream()) return flush_msm_on_side(true);
    int32_t rc = msm_g1_batched(ctx, bases.data(), scalars.data(), lens.data(), lens.size(), nullptr, outs.data(), pre.data());
    bases.clear(), scalars.clear(), lens.clear(), outs.clear(), pre.clear();
    entries = points = 0;
    return rc;
}
int32_t Deferred::run() {
    // the protocol phase (about a thousand short launches) ends here: what follows are the MSM launch sequences and the
    // leader rounds.  Hosts queue their next bulk host -> device copy behind this mark (scz_ctx_stream_wait_protocol_phase)
    if (!ctx->phase_mark) SCZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->phase_mark, cudaEventDisableTiming));
    SCZ_CUDA(ctx, cudaEventRecord(ctx->phase_mark, ctx->stream));
    using Fn = std::function<int32_t()>;
    // one leader stage over the first n1 continuations of `after` and the first n2 of `after2` (plus whatever the stage
    // itself appends to after2); continuations appended to `after` wait for the next flush
    auto stage = [&](size_t n1, size_t n2) -> int32_t {
        std::vector<Fn> now(std::make_move_iterator(after.begin()), std::make_move_iterator(after.begin() + n1));
        after.erase(after.begin(), after.begin() + n1);
        std::vector<Fn> now2(std::make_move_iterator(after2.begin()), std::make_move_iterator(after2.begin() + n2));
        after2.erase(after2.begin(), after2.begin() + n2);
        const size_t base2 = after2.size();
        for (auto &f : now) SCZ_TRY(f());
        SCZ_TRY(do_gathers());
        std::vector<Fn> g;
        g.swap(after_gather);
        for (auto &f : g) SCZ_TRY(f());
        SCZ_TRY(flush_closures());
        SCZ_TRY(do_scatters());
        for (auto &f : now2) SCZ_TRY(f());
        std::vector<Fn> appended(std::make_move_iterator(after2.begin() + base2), std::make_move_iterator(after2.end()));
        after2.erase(after2.begin() + base2, after2.end());
        for (auto &f : appended) SCZ_TRY(f());
        return SCZ_OK;
    };
    // SCZ_MSM_OVERLAP_CLOSURES=1 (dev knob, off): with an early-started sequence (flush_early) and more MSMs queued after
    // it, start the second sequence on the side stream BEFORE the leader rounds of the first one's results, so that their
    // closures (two 255-doubling chains per round, latency-bound: ~3.7 ms per 2^20 proof) run under its bucket
    // accumulation.  Bit-identical proofs (the continuations registered before the early start only read results of the
    // first sequence; every party splits at the same place), but measured SLOWER on B200: 155.7 vs 153.2 ms per proof --
    // the closures' 32-thread CTAs take 18 ms instead of 3.7 next to the bucket kernels and cost the second sequence
    // more SM time than the 3 ms they hide.
    static const bool overlap = [] { const char *e = getenv("SCZ_MSM_OVERLAP_CLOSURES"); return e && e[0] == '1'; }();
    if (overlap && early_pending && early_after > 0 && !lens.empty() && early_after <= after.size() && early_after2 <= after2.size()) {
        const size_t late1 = after.size() - early_after, late2 = after2.size() - early_after2;
        SCZ_TRY(join_early());
        SCZ_TRY(flush_msm_on_side(false));
        SCZ_TRY(stage(early_after, early_after2));
        SCZ_TRY(join_early());
        SCZ_TRY(stage(late1, late2));
    }
    early_after = early_after2 = 0;
    while (!lens.empty() || !after.empty() || !gathers.empty() || !after_gather.empty() || !pss_jobs.empty() ||
           !colsum_jobs.empty() || !scatters.empty() || !after2.empty() || early_pending) {
        SCZ_TRY(flush_msm());
        SCZ_TRY(stage(after.size(), after2.size()));
    }
    return SCZ_OK;
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_msm_g1_batched_dev(scz_ctx *h, const void *const *d_bases, const void *const *d_scalars, const size_t *lens,
                               size_t batch, void *d_out) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (batch && (!d_bases || !d_scalars || !lens || !d_out)) return h->c.fail(SCZ_ERR_BAD_ARG, "msm: null argument");
    return msm_g1_batched(&h->c, d_bases, d_scalars, lens, batch, d_out);
}

int32_t scz_msm_g1(scz_ctx *h, const void *bases, const uint8_t *inf_mask, size_t bases_len, const void *scalars,
                   size_t scalars_len, void *out_jac) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (!out_jac) return c->fail(SCZ_ERR_BAD_ARG, "msm: null output");
    if (bases_len != scalars_len)   // ark-ec returns Err(min len); the reference unwrap()s it (dmsm.rs:23)
        return c->fail(SCZ_ERR_LEN_MISMATCH, "msm: %zu bases vs %zu scalars", bases_len, scalars_len);
    size_t n = bases_len;
    if (n && (!bases || !scalars)) return c->fail(SCZ_ERR_BAD_ARG, "msm: null input");
    DevTmp d_b(c), d_s(c), d_m(c), d_o(c);
    SCZ_TRY(d_b.alloc(n * SCZ_G1_AFFINE_BYTES));
    SCZ_TRY(d_s.alloc(n * SCZ_FR_BYTES));
    SCZ_TRY(d_o.alloc(SCZ_G1_JAC_BYTES));
    if (n) {
        SCZ_CUDA(c, cudaMemcpyAsync(d_b.p, bases, n * SCZ_G1_AFFINE_BYTES, cudaMemcpyHostToDevice, c->stream));
        SCZ_CUDA(c, cudaMemcpyAsync(d_s.p, scalars, n * SCZ_FR_BYTES, cudaMemcpyHostToDevice, c->stream));
        if (inf_mask) {
            SCZ_TRY(d_m.alloc(n));
            SCZ_CUDA(c, cudaMemcpyAsync(d_m.p, inf_mask, n, cudaMemcpyHostToDevice, c->stream));
            SCZ_TRY(scz_g1_apply_inf_mask_dev(h, d_b.p, d_m.as<uint8_t>(), n));
        }
    }
    const void *bp = d_b.p, *sp = d_s.p;
    SCZ_TRY(msm_g1_batched(c, &bp, &sp, &n, 1, d_o.p));
    SCZ_CUDA(c, cudaMemcpyAsync(out_jac, d_o.p, SCZ_G1_JAC_BYTES, cudaMemcpyDeviceToHost, c->stream));
    SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    return SCZ_OK;
}

int32_t scz_msm_set_window(scz_ctx *h, uint32_t cbits) {
    scz::DeviceGuard dg__(h);
    if (!h || cbits > 20) return SCZ_ERR_BAD_ARG;
    h->c.msm_window_override = cbits;
    return SCZ_OK;
}
int32_t scz_msm_set_affine(scz_ctx *h, uint32_t mode, uint32_t levels, uint64_t slab_entries) {
    if (!h || mode > 2 || levels > 8) return SCZ_ERR_BAD_ARG;
    h->c.msm_affine_mode = mode;
    h->c.msm_affine_levels = levels;
    h->c.msm_affine_slab = slab_entries;
    return SCZ_OK;
}
int32_t scz_msm_affine_sequences(const scz_ctx *h, uint64_t *sequences) {
    if (!h || !sequences) return SCZ_ERR_BAD_ARG;
    *sequences = h->c.msm_cum_affine_sequences;
    return SCZ_OK;
}
int32_t scz_msm_use_precompute(scz_ctx *h, int32_t on) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    h->c.msm_no_precompute = on == 0;
    return SCZ_OK;
}
int32_t scz_msm_cum_stats(const scz_ctx *h, uint64_t *adds, uint64_t *pairs, uint64_t *sequences, uint64_t *segments) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (adds) *adds = h->c.msm_cum_adds;
    if (pairs) *pairs = h->c.msm_cum_pairs;
    if (sequences) *sequences = h->c.msm_cum_sequences;
    if (segments) *segments = h->c.msm_cum_segments;
    return SCZ_OK;
}
int32_t scz_msm_last_stats(const scz_ctx *h, uint64_t *adds, uint64_t *buckets, uint64_t *windows) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (adds) *adds = h->c.msm_bucket_adds;
    if (buckets) *buckets = h->c.msm_buckets;
    if (windows) *windows = h->c.msm_windows;
    return SCZ_OK;
}

}   // extern "C"
