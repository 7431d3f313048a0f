// Fr-table kernels of the path: product sumcheck rounds, PST open folds, the product tree,
// point-wise maps with batch inversion, fix_variable.
//
//   product sumcheck round   dist-primitive/src/dsumcheck.rs:37-85 (same loop body at :167-219, :227-279,
//                            :377-429, :452-504)
//   PST open fold            dist-primitive/src/dpoly_comm.rs:309-323 (= :337-351, :418-432)
//   product tree             dist-primitive/src/dacc_product.rs:30-39 / :374-381
//   fix_variable             dist-primitive/src/mle.rs:88-104
//   point-wise maps          hyperplonk/src/dhyperplonk.rs:233-238, 251-256, 326-339
//
// A table is a dense array of 32 B Fr elements (arkworks' Montgomery limbs); every thread moves whole
// elements with 128-bit loads and stores, consecutive threads touch consecutive elements.  The reference
// takes every challenge up-front (dsumcheck.rs:151), so a round's fold is fused with the evaluation of the
// SAME round's three sums: one pass reads (f_lo, f_hi, g_lo, g_hi) = 128 B per pair and writes the folded
// pair = 64 B, five Fr products in between.  a*(1-r) + b*r is computed as a + r*(b - a): the same field
// element with one product instead of two.
#include <stdlib.h>

#include "batch_inv.cuh"
#include "ctx.h"
#include "field.cuh"

namespace scz {

constexpr int PL_THREADS = 256;
constexpr uint32_t TAIL_PAIRS = 256;    // rounds with at most this many pairs finish inside one CTA (one pair per thread)

struct Fr3 {
    Fr a, b, c;
};

__device__ __forceinline__ Fr fr_shfl_down(const Fr &v, int delta) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], delta);
    return r;
}
// block-wide sum of three Fr accumulators; result valid in thread 0.  sh: 3 * (T/32) Fr
template <int T>
__device__ __forceinline__ void block_sum3(Fr &s0, Fr &s1, Fr &s2, Fr *sh) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s0 = fp_add(s0, fr_shfl_down(s0, d));
        s1 = fp_add(s1, fr_shfl_down(s1, d));
        s2 = fp_add(s2, fr_shfl_down(s2, d));
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        sh[3 * wid] = s0;
        sh[3 * wid + 1] = s1;
        sh[3 * wid + 2] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < T / 32; w++) {
            s0 = fp_add(s0, sh[3 * w]);
            s1 = fp_add(s1, sh[3 * w + 1]);
            s2 = fp_add(s2, sh[3 * w + 2]);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void product_pair(const Fr &f0, const Fr &f1, const Fr &g0, const Fr &g1, const Fr &r,
                                             Fr &s0, Fr &s1, Fr &s2, Fr &fo, Fr &go) {
    s0 = fp_add(s0, fp_mul(f0, g0));
    s1 = fp_add(s1, fp_mul(f1, g1));
    Fr df = fp_sub(f1, f0), dg = fp_sub(g1, g0);
    s2 = fp_add(s2, fp_mul(fp_add(f1, df), fp_add(g1, dg)));   // (2 f1 - f0)(2 g1 - g0)
    fo = fp_add(f0, fp_mul(r, df));                             // f0 (1 - r) + f1 r
    go = fp_add(g0, fp_mul(r, dg));
}

// One round over h pairs, big rounds.  The three round sums are sums of products: a thread adds the plain 512-bit
// products of its pairs into three 17-limb accumulators and reduces each ONCE at the end (field.cuh fp_mul_acc_wide:
// 64 instead of 128 wide multiplies per term -- 436 instead of 610 multiply-pipe instructions per pair with the two fold
// products).  One CTA per SM, every thread strides over h with the next pair's four loads in flight while it works on
// the current one (228 registers; measured against 2 x 256 and 3 x 128 threads per SM without the prefetch: 1.38 vs
// 1.45 - 1.48 ms for all rounds of a 2^24-entry table).  Every CTA leaves its partial triple in `partial`; the last
// CTA to finish (ticket) adds them up and writes the round message.
constexpr int SC_THREADS = 256;
__global__ void __launch_bounds__(SC_THREADS, 1) k_sumcheck_round(const void *f_in, const void *g_in, void *f_out, void *g_out,
                                                                  uint32_t h, const void *challenge, Fr3 *partial,
                                                                  uint32_t *ticket, Fr3 *out) {
    constexpr int PL_THREADS = SC_THREADS;
    __shared__ Fr sh[3 * PL_THREADS / 32];
    __shared__ bool last;
    const Fr r = fp_load<FrP>(challenge, 0);
    uint32_t a0[17], a1[17], a2[17];
#pragma unroll
    for (int k = 0; k < 17; k++) a0[k] = a1[k] = a2[k] = 0;
    const uint32_t stride = gridDim.x * PL_THREADS;
    uint32_t i = blockIdx.x * PL_THREADS + threadIdx.x;
    Fr nf0, nf1, ng0, ng1;
    if (i < h) {
        nf0 = fp_load_rw<FrP>(f_in, i), nf1 = fp_load_rw<FrP>(f_in, (size_t)h + i);
        ng0 = fp_load_rw<FrP>(g_in, i), ng1 = fp_load_rw<FrP>(g_in, (size_t)h + i);
    }
    for (; i < h; i += stride) {
        const Fr f0 = nf0, f1 = nf1, g0 = ng0, g1 = ng1;
        const uint32_t n = i + stride < h ? i + stride : i;   // the next pair (the last iteration re-reads its own)
        nf0 = fp_load_rw<FrP>(f_in, n), nf1 = fp_load_rw<FrP>(f_in, (size_t)h + n);
        ng0 = fp_load_rw<FrP>(g_in, n), ng1 = fp_load_rw<FrP>(g_in, (size_t)h + n);
        fp_mul_acc_wide<FrP>(a0, f0, g0);
        fp_mul_acc_wide<FrP>(a1, f1, g1);
        Fr df = fp_sub(f1, f0), dg = fp_sub(g1, g0);
        fp_store<FrP>(f_out, i, fp_add(f0, fp_mul(r, df)));        // f0 (1 - r) + f1 r
        fp_store<FrP>(g_out, i, fp_add(g0, fp_mul(r, dg)));
        fp_mul_acc_wide<FrP>(a2, fp_add(f1, df), fp_add(g1, dg));   // (2 f1 - f0)(2 g1 - g0)
    }
    Fr s0 = fp_acc_wide_reduce<FrP>(a0), s1 = fp_acc_wide_reduce<FrP>(a1), s2 = fp_acc_wide_reduce<FrP>(a2);
    block_sum3<PL_THREADS>(s0, s1, s2, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x].a = s0;
        partial[blockIdx.x].b = s1;
        partial[blockIdx.x].c = s2;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    s0 = Fr::zero(), s1 = Fr::zero(), s2 = Fr::zero();
    for (uint32_t b = threadIdx.x; b < gridDim.x; b += PL_THREADS) {
        const uint4 *p = reinterpret_cast<const uint4 *>(&partial[b]);   // written by other CTAs: read through L2
        Fr3 t;
        uint32_t *w = &t.a.l[0];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            uint4 v = __ldcg(p + k);
            w[4 * k] = v.x, w[4 * k + 1] = v.y, w[4 * k + 2] = v.z, w[4 * k + 3] = v.w;
        }
        s0 = fp_add(s0, t.a);
        s1 = fp_add(s1, t.b);
        s2 = fp_add(s2, t.c);
    }
    block_sum3<PL_THREADS>(s0, s1, s2, sh);
    if (threadIdx.x == 0) {
        out->a = s0;
        out->b = s1;
        out->c = s2;
        *ticket = 0;   // ready for the next round
    }
}

// The same round with one Montgomery product per term (product_pair): for rounds with only a few pairs per thread,
// where the three accumulator reductions at the end of the lazy kernel would cost more than they save.
__global__ void __launch_bounds__(PL_THREADS) k_sumcheck_round_direct(const void *f_in, const void *g_in, void *f_out,
                                                                       void *g_out, uint32_t h, const void *challenge,
                                                                       Fr3 *partial, uint32_t *ticket, Fr3 *out) {
    __shared__ Fr sh[3 * PL_THREADS / 32];
    __shared__ bool last;
    const Fr r = fp_load<FrP>(challenge, 0);
    Fr s0 = Fr::zero(), s1 = Fr::zero(), s2 = Fr::zero();
    for (uint32_t i = blockIdx.x * PL_THREADS + threadIdx.x; i < h; i += gridDim.x * PL_THREADS) {
        Fr f0 = fp_load_rw<FrP>(f_in, i), f1 = fp_load_rw<FrP>(f_in, (size_t)h + i);
        Fr g0 = fp_load_rw<FrP>(g_in, i), g1 = fp_load_rw<FrP>(g_in, (size_t)h + i);
        Fr fo, go;
        product_pair(f0, f1, g0, g1, r, s0, s1, s2, fo, go);
        fp_store<FrP>(f_out, i, fo);
        fp_store<FrP>(g_out, i, go);
    }
    block_sum3<PL_THREADS>(s0, s1, s2, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x].a = s0;
        partial[blockIdx.x].b = s1;
        partial[blockIdx.x].c = s2;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    s0 = Fr::zero(), s1 = Fr::zero(), s2 = Fr::zero();
    for (uint32_t b = threadIdx.x; b < gridDim.x; b += PL_THREADS) {
        const uint4 *p = reinterpret_cast<const uint4 *>(&partial[b]);   // written by other CTAs: read through L2
        Fr3 t;
        uint32_t *w = &t.a.l[0];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            uint4 v = __ldcg(p + k);
            w[4 * k] = v.x, w[4 * k + 1] = v.y, w[4 * k + 2] = v.z, w[4 * k + 3] = v.w;
        }
        s0 = fp_add(s0, t.a);
        s1 = fp_add(s1, t.b);
        s2 = fp_add(s2, t.c);
    }
    block_sum3<PL_THREADS>(s0, s1, s2, sh);
    if (threadIdx.x == 0) {
        out->a = s0;
        out->b = s1;
        out->c = s2;
        *ticket = 0;
    }
}

// All remaining rounds (h <= TAIL_PAIRS pairs in the first of them) inside one CTA, in place on f/g.
// challenge points at the first of these rounds' challenges; out at their first message.
__global__ void __launch_bounds__(PL_THREADS) k_sumcheck_tail(void *f, void *g, uint32_t h, const void *challenge,
                                                               Fr3 *out) {
    __shared__ Fr sh[3 * PL_THREADS / 32];
    for (uint32_t round = 0; h >= 1; h >>= 1, round++) {
        const Fr r = fp_load<FrP>(challenge, round);
        Fr s0 = Fr::zero(), s1 = Fr::zero(), s2 = Fr::zero();
        Fr fo[TAIL_PAIRS / PL_THREADS], go[TAIL_PAIRS / PL_THREADS];
#pragma unroll
        for (uint32_t k = 0; k < TAIL_PAIRS / PL_THREADS; k++) {
            uint32_t i = k * PL_THREADS + threadIdx.x;
            if (i < h) {
                Fr f0 = fp_load_rw<FrP>(f, i), f1 = fp_load_rw<FrP>(f, (size_t)h + i);
                Fr g0 = fp_load_rw<FrP>(g, i), g1 = fp_load_rw<FrP>(g, (size_t)h + i);
                product_pair(f0, f1, g0, g1, r, s0, s1, s2, fo[k], go[k]);
            }
        }
        __syncthreads();   // every read of this round is done before any slot is overwritten
#pragma unroll
        for (uint32_t k = 0; k < TAIL_PAIRS / PL_THREADS; k++) {
            uint32_t i = k * PL_THREADS + threadIdx.x;
            if (i < h) {
                fp_store<FrP>(f, i, fo[k]);
                fp_store<FrP>(g, i, go[k]);
            }
        }
        block_sum3<PL_THREADS>(s0, s1, s2, sh);
        if (threadIdx.x == 0) {
            out[round].a = s0;
            out[round].b = s1;
            out[round].c = s2;
        }
        __syncthreads();
    }
}

// ---- single-MLE sumcheck (dsumcheck.rs:6-26; the same loop at :107-121, :128-141, :303-316, :335-347): the round message
// is (sum lo, sum hi), the fold is lo + r (hi - lo).  One Fr product per pair against 64 B read + 32 B written.
struct Fr2 {
    Fr a, b;
};
__global__ void __launch_bounds__(PL_THREADS) k_sum_round(const void *f_in, void *f_out, uint32_t h, const void *challenge,
                                                           Fr3 *partial, uint32_t *ticket, Fr2 *out) {
    __shared__ Fr sh[3 * PL_THREADS / 32];
    __shared__ bool last;
    const Fr r = fp_load<FrP>(challenge, 0);
    Fr s0 = Fr::zero(), s1 = Fr::zero(), s2 = Fr::zero();
    for (uint32_t i = blockIdx.x * PL_THREADS + threadIdx.x; i < h; i += gridDim.x * PL_THREADS) {
        Fr f0 = fp_load_rw<FrP>(f_in, i), f1 = fp_load_rw<FrP>(f_in, (size_t)h + i);
        s0 = fp_add(s0, f0);
        s1 = fp_add(s1, f1);
        fp_store<FrP>(f_out, i, fp_add(f0, fp_mul(r, fp_sub(f1, f0))));
    }
    block_sum3<PL_THREADS>(s0, s1, s2, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x].a = s0;
        partial[blockIdx.x].b = s1;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    s0 = Fr::zero(), s1 = Fr::zero();
    for (uint32_t b = threadIdx.x; b < gridDim.x; b += PL_THREADS) {
        const uint4 *p = reinterpret_cast<const uint4 *>(&partial[b]);   // written by other CTAs: read through L2
        Fr2 t;
        uint32_t *w = &t.a.l[0];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint4 v = __ldcg(p + k);
            w[4 * k] = v.x, w[4 * k + 1] = v.y, w[4 * k + 2] = v.z, w[4 * k + 3] = v.w;
        }
        s0 = fp_add(s0, t.a);
        s1 = fp_add(s1, t.b);
    }
    block_sum3<PL_THREADS>(s0, s1, s2, sh);
    if (threadIdx.x == 0) {
        out->a = s0;
        out->b = s1;
        *ticket = 0;
    }
}
// R rounds of the single-MLE sumcheck in ONE pass (every challenge is known up-front, dsumcheck.rs:151): a thread
// holds the 2^R entries i + m w (w = len >> R) that fold into entry i, adds them to the (lo, hi) sums of each of the R
// rounds as it folds them down.  Traffic per folded entry: 2^R * 32 B read + 32 B written instead of 96 B per pair per
// round (R = 3: 288 B instead of 672 B), and a third of the launches -- this chain of halving passes is bandwidth- and
// launch-bound, not compute-bound (one product per pair).
template <int R, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) k_sum_rounds(const void *f_in, void *f_out, uint32_t w, const void *challenge,
                                                        Fr *partial, uint32_t *ticket, Fr2 *out) {
    constexpr int PL_THREADS = T;
    constexpr int M = 1 << R, NV = 2 * R;
    __shared__ Fr sh[NV * PL_THREADS / 32];
    __shared__ bool last;
    // (measured, not kept: the round sums as unreduced 9-limb integers reduced once per thread -- 9 add-with-carry
    // instead of fp_add's 26 instructions per term, but 6 more registers at the 168-register cap: 0.199 / 0.466 ms
    // instead of 0.179 / 0.440 ms for a whole call at 2^22 / 2^24 entries)
    Fr r[R], s[NV];
#pragma unroll
    for (int t = 0; t < R; t++) r[t] = fp_load<FrP>(challenge, t);
#pragma unroll
    for (int k = 0; k < NV; k++) s[k] = Fr::zero();
    for (uint32_t i = blockIdx.x * PL_THREADS + threadIdx.x; i < w; i += gridDim.x * PL_THREADS) {
        Fr v[M];
#pragma unroll
        for (int m = 0; m < M; m++) v[m] = fp_load_rw<FrP>(f_in, (size_t)m * w + i);
#pragma unroll
        for (int t = 0; t < R; t++) {
            constexpr int dummy = 0;
            (void)dummy;
            const int half = M >> (t + 1);
#pragma unroll
            for (int m = 0; m < M / 2; m++) {
                if (m < half) {
                    s[2 * t] = fp_add(s[2 * t], v[m]);
                    s[2 * t + 1] = fp_add(s[2 * t + 1], v[m + half]);
                    v[m] = fp_add(v[m], fp_mul(r[t], fp_sub(v[m + half], v[m])));
                }
            }
        }
        fp_store<FrP>(f_out, i, v[0]);
    }
    // block-wide sums of the NV accumulators
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s[k] = fp_add(s[k], fr_shfl_down(s[k], d));
        if (lane == 0) sh[k * (PL_THREADS / 32) + wid] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        Fr acc = sh[threadIdx.x * (PL_THREADS / 32)];
        for (int q = 1; q < PL_THREADS / 32; q++) acc = fp_add(acc, sh[threadIdx.x * (PL_THREADS / 32) + q]);
        fp_store<FrP>(partial, (size_t)blockIdx.x * NV + threadIdx.x, acc);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last CTA adds the partials: one warp per accumulator
    for (int acc_k = wid; acc_k < NV; acc_k += PL_THREADS / 32) {
        Fr acc = Fr::zero();
        for (uint32_t b = lane; b < gridDim.x; b += 32) {
            const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(partial) + ((size_t)b * NV + acc_k) * 32);
            Fr t;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                uint4 x = __ldcg(p + q);   // written by other CTAs: read through L2
                t.l[4 * q] = x.x, t.l[4 * q + 1] = x.y, t.l[4 * q + 2] = x.z, t.l[4 * q + 3] = x.w;
            }
            acc = fp_add(acc, t);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc = fp_add(acc, fr_shfl_down(acc, d));
        if (lane == 0) {
            Fr *o = reinterpret_cast<Fr *>(out);   // round t: (sum lo, sum hi) = accumulators 2t, 2t + 1
            o[acc_k] = acc;
        }
    }
    if (threadIdx.x == 0) *ticket = 0;
}

// all remaining rounds (h <= TAIL_PAIRS) in one CTA, in place on f
__global__ void __launch_bounds__(PL_THREADS) k_sum_tail(void *f, uint32_t h, const void *challenge, Fr2 *out) {
    __shared__ Fr sh[3 * PL_THREADS / 32];
    for (uint32_t round = 0; h >= 1; h >>= 1, round++) {
        const Fr r = fp_load<FrP>(challenge, round);
        Fr s0 = Fr::zero(), s1 = Fr::zero(), s2 = Fr::zero(), fo = Fr::zero();
        uint32_t i = threadIdx.x;
        if (i < h) {
            Fr f0 = fp_load_rw<FrP>(f, i), f1 = fp_load_rw<FrP>(f, (size_t)h + i);
            s0 = f0;
            s1 = f1;
            fo = fp_add(f0, fp_mul(r, fp_sub(f1, f0)));
        }
        __syncthreads();
        if (i < h) fp_store<FrP>(f, i, fo);
        block_sum3<PL_THREADS>(s0, s1, s2, sh);
        if (threadIdx.x == 0) {
            out[round].a = s0;
            out[round].b = s1;
        }
        __syncthreads();
    }
}

// PST open fold round: q[j] = hi - lo ; cur[j] = lo + u * q[j]   (= (1-u) lo + u hi)
__global__ void __launch_bounds__(PL_THREADS) k_open_fold(const void *in, void *cur_out, void *q, uint32_t h,
                                                           const void *point) {
    const Fr u = fp_load<FrP>(point, 0);
    for (uint32_t i = blockIdx.x * PL_THREADS + threadIdx.x; i < h; i += gridDim.x * PL_THREADS) {
        Fr lo = fp_load_rw<FrP>(in, i), hi = fp_load_rw<FrP>(in, (size_t)h + i);
        Fr d = fp_sub(hi, lo);
        fp_store<FrP>(q, i, d);
        fp_store<FrP>(cur_out, i, fp_add(lo, fp_mul(u, d)));
    }
}
// tail: all remaining rounds in one CTA; q arena is laid out round after round (h, h/2, ..., 1 entries)
__global__ void __launch_bounds__(PL_THREADS) k_open_fold_tail(void *cur, void *q, uint32_t h, const void *point) {
    size_t qoff = 0;
    for (uint32_t round = 0; h >= 1; qoff += h, h >>= 1, round++) {
        const Fr u = fp_load<FrP>(point, round);
        Fr keep[TAIL_PAIRS / PL_THREADS];
#pragma unroll
        for (uint32_t k = 0; k < TAIL_PAIRS / PL_THREADS; k++) {
            uint32_t i = k * PL_THREADS + threadIdx.x;
            if (i < h) {
                Fr lo = fp_load_rw<FrP>(cur, i), hi = fp_load_rw<FrP>(cur, (size_t)h + i);
                Fr d = fp_sub(hi, lo);
                fp_store<FrP>(q, qoff + i, d);
                keep[k] = fp_add(lo, fp_mul(u, d));
            }
        }
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < TAIL_PAIRS / PL_THREADS; k++) {
            uint32_t i = k * PL_THREADS + threadIdx.x;
            if (i < h) fp_store<FrP>(cur, i, keep[k]);
        }
        __syncthreads();
    }
}

// fix_variable round: out[j] = lo + p * (hi - lo)
__global__ void __launch_bounds__(PL_THREADS) k_fix_variable(const void *in, void *out, uint32_t h, const void *point) {
    const Fr u = fp_load<FrP>(point, 0);
    for (uint32_t i = blockIdx.x * PL_THREADS + threadIdx.x; i < h; i += gridDim.x * PL_THREADS) {
        Fr lo = fp_load_rw<FrP>(in, i), hi = fp_load_rw<FrP>(in, (size_t)h + i);
        fp_store<FrP>(out, i, fp_add(lo, fp_mul(u, fp_sub(hi, lo))));
    }
}

// product tree level: tree[dst + k] = tree[src + 2k] * tree[src + 2k + 1], k < cnt
__global__ void __launch_bounds__(PL_THREADS) k_tree_level(void *tree, size_t src, size_t dst, uint32_t cnt) {
    uint32_t k = blockIdx.x * PL_THREADS + threadIdx.x;
    if (k >= cnt) return;
    fp_store<FrP>(tree, dst + k, fp_mul(fp_load_rw<FrP>(tree, src + 2 * (size_t)k), fp_load_rw<FrP>(tree, src + 2 * (size_t)k + 1)));
}
// the levels with at most PL_THREADS products each, in one CTA; finally tree[2m-1] = 0 (dacc_product.rs:39,381)
__global__ void __launch_bounds__(PL_THREADS) k_tree_tail(void *tree, size_t src, size_t dst, uint32_t cnt, size_t last) {
    for (; cnt >= 1; cnt >>= 1) {
        if (threadIdx.x < cnt)
            fp_store<FrP>(tree, dst + threadIdx.x,
                          fp_mul(fp_load_rw<FrP>(tree, src + 2 * (size_t)threadIdx.x),
                                 fp_load_rw<FrP>(tree, src + 2 * (size_t)threadIdx.x + 1)));
        __syncthreads();
        src = dst;
        dst += cnt;
    }
    if (threadIdx.x == 0) fp_store<FrP>(tree, last, Fr::zero());
}

// point-wise: out = a*ca + b*cb + cc with constants from a small device array k = (ca, cb, cc);
// mode 0: out = a + b ; 1: out = b - a ; 2: out = a + k0*b + k1   (num / den of dhyperplonk.rs:326-337)
template <int MODE>
__global__ void __launch_bounds__(PL_THREADS) k_pointwise(const void *a, const void *b, const void *k, void *out,
                                                           size_t n) {
    size_t i = blockIdx.x * (size_t)PL_THREADS + threadIdx.x;
    if (i >= n) return;
    Fr x = fp_load_rw<FrP>(a, i), y = fp_load_rw<FrP>(b, i), r;
    if (MODE == 0) r = fp_add(x, y);
    else if (MODE == 1) r = fp_sub(y, x);
    else r = fp_add(fp_add(x, fp_mul(fp_load<FrP>(k, 0), y)), fp_load<FrP>(k, 1));
    fp_store<FrP>(out, i, r);
}

// out[i] = num[i] / den[i] (dhyperplonk.rs:338-339) with ONE field inversion for the whole table (Montgomery's trick as
// a device-wide product tree, batch_inv.cuh): phase 1 writes per-thread running products of the denominators and the
// thread totals, the tree turns the totals into their inverses, phase 2 walks each thread's elements backwards.  Four
// products per element, 192 B of traffic against 96 B algorithmic; the earlier version ran one 255-bit Fermat chain per
// 1024 elements on a single lane of each CTA, which serialised the kernel (3.6 ms for 2^22 elements, now ~0.2 ms).
// den = 0: arkworks panics (Field::div); here the element comes out 0, the rest of the batch stays right and the ctx's
// SCZ_STATUS_DIV_BY_ZERO bit is raised.
constexpr int DIV_THREADS = 256;
constexpr int DIV_PER = 4;      // elements per thread, strided by the CTA (coalesced)
__global__ void __launch_bounds__(DIV_THREADS) k_div_phase1(const void *den, size_t n, void *pre, void *tot, uint32_t *status) {
    const size_t cta_base = (size_t)blockIdx.x * (DIV_THREADS * DIV_PER);
    Fr run = Fr::one();
    bool zero = false;
#pragma unroll
    for (int j = 0; j < DIV_PER; j++) {
        size_t e = cta_base + (size_t)j * DIV_THREADS + threadIdx.x;
        if (e < n) {
            Fr d = fp_load_rw<FrP>(den, e);
            if (d.is_zero()) zero = true;                  // keeps the batch invertible; the element is fixed up in phase 2
            else run = fp_mul(run, d);
            fp_store<FrP>(pre, e, run);
        }
    }
    fp_store<FrP>(tot, (size_t)blockIdx.x * DIV_THREADS + threadIdx.x, run);
    if (zero) atomicOr(status, SCZ_STATUS_DIV_BY_ZERO);
}
__global__ void __launch_bounds__(DIV_THREADS) k_div_phase2(const void *num, const void *den, size_t n, const void *pre,
                                                             const void *inv_tot, void *out) {
    const size_t cta_base = (size_t)blockIdx.x * (DIV_THREADS * DIV_PER);
    Fr I = fp_load_rw<FrP>(inv_tot, (size_t)blockIdx.x * DIV_THREADS + threadIdx.x);
#pragma unroll
    for (int j = DIV_PER - 1; j >= 0; j--) {
        size_t e = cta_base + (size_t)j * DIV_THREADS + threadIdx.x;
        if (e >= n) continue;
        Fr d = fp_load_rw<FrP>(den, e);
        if (d.is_zero()) {
            fp_store<FrP>(out, e, Fr::zero());
            continue;
        }
        Fr dinv = j ? fp_mul(I, fp_load_rw<FrP>(pre, e - DIV_THREADS)) : I;
        if (j) I = fp_mul(I, d);
        fp_store<FrP>(out, e, fp_mul(fp_load_rw<FrP>(num, e), dinv));
    }
}

static uint32_t grid_for(Ctx *c, size_t items, int threads = PL_THREADS, int ctas_per_sm = 8) {
    size_t want = (items + threads - 1) / threads;
    size_t cap = (size_t)c->sm_count * ctas_per_sm;   // a multiple of the SM count; grid-stride loops cover the rest
    return (uint32_t)(want < cap ? (want ? want : 1) : cap);
}

// n = log2(len) rounds of the product sumcheck on device tables; writes n triples to d_out and the two
// fully folded values (f, g) to d_last.  The inputs are left untouched (the reference clones, dsumcheck.rs:160).
int32_t sumcheck_product_rounds(Ctx *ctx, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                                void *d_out, void *d_last) {
    if (len == 0 || (len & (len - 1))) return ctx->fail(SCZ_ERR_NOT_POW2, "sumcheck: table length %zu is not a power of two", len);
    if (len >= (1ull << 32)) return ctx->fail(SCZ_ERR_BAD_ARG, "sumcheck: table too long");
    ProfScope ps(ctx, SCZ_K_SUMCHECK);
    cudaStream_t st = ctx->stream;
    if (len == 1) {
        SCZ_CUDA(ctx, cudaMemcpyAsync(d_last, d_f, 32, cudaMemcpyDeviceToDevice, st));
        SCZ_CUDA(ctx, cudaMemcpyAsync((char *)d_last + 32, d_g, 32, cudaMemcpyDeviceToDevice, st));
        return SCZ_OK;
    }
    size_t h = len / 2;
    if (h <= TAIL_PAIRS) {   // short table: everything in one CTA, in place on a private copy
        DevTmp ff(ctx), gg(ctx);
        SCZ_TRY(ff.alloc(len * 32));
        SCZ_TRY(gg.alloc(len * 32));
        SCZ_CUDA(ctx, cudaMemcpyAsync(ff.p, d_f, len * 32, cudaMemcpyDeviceToDevice, st));
        SCZ_CUDA(ctx, cudaMemcpyAsync(gg.p, d_g, len * 32, cudaMemcpyDeviceToDevice, st));
        k_sumcheck_tail<<<1, PL_THREADS, 0, st>>>(ff.p, gg.p, (uint32_t)h, d_challenge, reinterpret_cast<Fr3 *>(d_out));
        SCZ_LAUNCH_CHECK(ctx);
        SCZ_CUDA(ctx, cudaMemcpyAsync(d_last, ff.p, 32, cudaMemcpyDeviceToDevice, st));
        SCZ_CUDA(ctx, cudaMemcpyAsync((char *)d_last + 32, gg.p, 32, cudaMemcpyDeviceToDevice, st));
        return SCZ_OK;
    }
    DevTmp tf(ctx), tg(ctx), partial(ctx), ticket(ctx);
    SCZ_TRY(tf.alloc(h * 32));
    SCZ_TRY(tg.alloc(h * 32));
    SCZ_TRY(partial.alloc((size_t)ctx->sm_count * 16 * sizeof(Fr3)));
    SCZ_TRY(ticket.alloc(4));
    SCZ_CUDA(ctx, cudaMemsetAsync(ticket.p, 0, 4, st));
    const void *fi = d_f, *gi = d_g;
    size_t round = 0;
    // the lazy kernel from ~4 pairs per thread of its one-CTA-per-SM grid, the direct one below (measured on B200)
    const size_t lazy_from = (size_t)4 * ctx->sm_count * SC_THREADS;
    while (h > TAIL_PAIRS) {
        const void *ch = (const char *)d_challenge + round * 32;
        Fr3 *o = reinterpret_cast<Fr3 *>(d_out) + round;
        if (h >= lazy_from)
            k_sumcheck_round<<<ctx->sm_count, SC_THREADS, 0, st>>>(fi, gi, tf.p, tg.p, (uint32_t)h, ch, partial.as<Fr3>(), ticket.as<uint32_t>(), o);
        else
            k_sumcheck_round_direct<<<grid_for(ctx, h), PL_THREADS, 0, st>>>(fi, gi, tf.p, tg.p, (uint32_t)h, ch, partial.as<Fr3>(), ticket.as<uint32_t>(), o);
        SCZ_LAUNCH_CHECK(ctx);
        fi = tf.p;
        gi = tg.p;
        h >>= 1;
        round++;
    }
    k_sumcheck_tail<<<1, PL_THREADS, 0, st>>>(tf.p, tg.p, (uint32_t)h, (const char *)d_challenge + round * 32,
                                              reinterpret_cast<Fr3 *>(d_out) + round);
    SCZ_LAUNCH_CHECK(ctx);
    SCZ_CUDA(ctx, cudaMemcpyAsync(d_last, tf.p, 32, cudaMemcpyDeviceToDevice, st));
    SCZ_CUDA(ctx, cudaMemcpyAsync((char *)d_last + 32, tg.p, 32, cudaMemcpyDeviceToDevice, st));
    return SCZ_OK;
}

// n = log2(len) rounds of the single-MLE sumcheck: n pairs to d_out, the fully folded value to d_last (32 B)
int32_t sumcheck_rounds(Ctx *ctx, const void *d_f, size_t len, const void *d_challenge, void *d_out, void *d_last) {
    if (len == 0 || (len & (len - 1))) return ctx->fail(SCZ_ERR_NOT_POW2, "sumcheck: table length %zu is not a power of two", len);
    if (len >= (1ull << 32)) return ctx->fail(SCZ_ERR_BAD_ARG, "sumcheck: table too long");
    static_assert(TAIL_PAIRS == PL_THREADS, "k_sum_tail handles one pair per thread");
    ProfScope ps(ctx, SCZ_K_SUMCHECK);
    cudaStream_t st = ctx->stream;
    if (len == 1) {
        SCZ_CUDA(ctx, cudaMemcpyAsync(d_last, d_f, 32, cudaMemcpyDeviceToDevice, st));
        return SCZ_OK;
    }
    size_t h = len / 2, round = 0;
    DevTmp tf(ctx), partial(ctx), ticket(ctx);
    SCZ_TRY(tf.alloc((h > TAIL_PAIRS ? h : len) * 32));
    SCZ_TRY(partial.alloc((size_t)ctx->sm_count * 16 * 6 * sizeof(Fr)));

    SCZ_TRY(ticket.alloc(4));
    SCZ_CUDA(ctx, cudaMemsetAsync(ticket.p, 0, 4, st));
    const void *fi = d_f;
    while (h > TAIL_PAIRS) {
        // up to three rounds per pass while more than TAIL_PAIRS pairs remain afterwards
        int R = 1;   // (measured at 2^24 entries: 0.59 / 0.54 / 0.55 ms with at most 1 / 2 / 3 rounds per pass)
        while (R < 3 && (h >> R) > TAIL_PAIRS) R++;
        const uint32_t w = (uint32_t)((2 * h) >> R);
        const void *ch = (const char *)d_challenge + round * 32;
        Fr2 *o = reinterpret_cast<Fr2 *>(d_out) + round;
        Fr *pa = partial.as<Fr>();
        uint32_t *tk = ticket.as<uint32_t>();
        // R = 3 runs with 128-thread CTAs, three per SM (168 registers): 0.44 ms for a 2^24-entry table against 0.51 ms
        // with one 256-thread CTA per SM
        if (R == 3) k_sum_rounds<3, 128, 3><<<grid_for(ctx, w, 128, 3), 128, 0, st>>>(fi, tf.p, w, ch, pa, tk, o);
        else if (R == 2) k_sum_rounds<2, 256, 2><<<grid_for(ctx, w, 256, 2), 256, 0, st>>>(fi, tf.p, w, ch, pa, tk, o);
        else k_sum_round<<<grid_for(ctx, h), PL_THREADS, 0, st>>>(fi, tf.p, (uint32_t)h, ch, partial.as<Fr3>(), tk, o);
        SCZ_LAUNCH_CHECK(ctx);
        fi = tf.p;
        h >>= R;
        round += R;
    }
    if (fi == d_f) SCZ_CUDA(ctx, cudaMemcpyAsync(tf.p, d_f, len * 32, cudaMemcpyDeviceToDevice, st));
    k_sum_tail<<<1, PL_THREADS, 0, st>>>(tf.p, (uint32_t)h, (const char *)d_challenge + round * 32,
                                         reinterpret_cast<Fr2 *>(d_out) + round);
    SCZ_LAUNCH_CHECK(ctx);
    SCZ_CUDA(ctx, cudaMemcpyAsync(d_last, tf.p, 32, cudaMemcpyDeviceToDevice, st));
    return SCZ_OK;
}

// n fold rounds of a PST opening: quotient tables q_0 .. q_{n-1} (len/2, len/4, ..., 1 entries) packed one
// after the other into d_q (len - 1 entries), the evaluation into d_value.
int32_t open_fold_rounds(Ctx *ctx, const void *d_peval, size_t len, const void *d_point, void *d_q, void *d_value) {
    if (len == 0 || (len & (len - 1))) return ctx->fail(SCZ_ERR_NOT_POW2, "open: table length %zu is not a power of two", len);
    if (len >= (1ull << 32)) return ctx->fail(SCZ_ERR_BAD_ARG, "open: table too long");
    ProfScope ps(ctx, SCZ_K_OPEN_FOLD);
    cudaStream_t st = ctx->stream;
    if (len == 1) {
        SCZ_CUDA(ctx, cudaMemcpyAsync(d_value, d_peval, 32, cudaMemcpyDeviceToDevice, st));
        return SCZ_OK;
    }
    size_t h = len / 2, qoff = 0, round = 0;
    DevTmp cur(ctx);
    SCZ_TRY(cur.alloc((h > TAIL_PAIRS ? h : len) * 32));
    const void *in = d_peval;
    while (h > TAIL_PAIRS) {
        k_open_fold<<<grid_for(ctx, h), PL_THREADS, 0, st>>>(in, cur.p, (char *)d_q + qoff * 32, (uint32_t)h,
                                                             (const char *)d_point + round * 32);
        SCZ_LAUNCH_CHECK(ctx);
        in = cur.p;
        qoff += h;
        h >>= 1;
        round++;
    }
    if (in == d_peval) SCZ_CUDA(ctx, cudaMemcpyAsync(cur.p, d_peval, len * 32, cudaMemcpyDeviceToDevice, st));
    k_open_fold_tail<<<1, PL_THREADS, 0, st>>>(cur.p, (char *)d_q + qoff * 32, (uint32_t)h,
                                               (const char *)d_point + round * 32);
    SCZ_LAUNCH_CHECK(ctx);
    SCZ_CUDA(ctx, cudaMemcpyAsync(d_value, cur.p, 32, cudaMemcpyDeviceToDevice, st));
    return SCZ_OK;
}

int32_t fix_variable_rounds(Ctx *ctx, const void *d_evals, size_t len, const void *d_points, size_t npoints, void *d_out) {
    if (len == 0 || (len & (len - 1))) return ctx->fail(SCZ_ERR_NOT_POW2, "fix_variable: length %zu is not a power of two", len);
    ProfScope ps(ctx, SCZ_K_POINTWISE);
    cudaStream_t st = ctx->stream;
    size_t n = 0;
    while (((size_t)1 << n) < len) n++;
    size_t k = npoints < n ? npoints : n;
    if (k == 0) {
        SCZ_CUDA(ctx, cudaMemcpyAsync(d_out, d_evals, len * 32, cudaMemcpyDeviceToDevice, st));
        return SCZ_OK;
    }
    DevTmp a(ctx), b(ctx);
    SCZ_TRY(a.alloc(len / 2 * 32));
    SCZ_TRY(b.alloc(len / 2 * 32));
    const void *in = d_evals;
    size_t h = len / 2;
    for (size_t i = 0; i < k; i++, h >>= 1) {
        void *out = i + 1 == k ? d_out : (i & 1 ? b.p : a.p);
        k_fix_variable<<<grid_for(ctx, h), PL_THREADS, 0, st>>>(in, out, (uint32_t)h, (const char *)d_points + i * 32);
        SCZ_LAUNCH_CHECK(ctx);
        in = out;
    }
    return SCZ_OK;
}

// tree (2m entries): [0, m) = x, [m, 2m-1) products level by level, [2m-1] = 0   (dacc_product.rs:30-39)
int32_t acc_product_tree(Ctx *ctx, const void *d_x, size_t m, void *d_tree) {
    if (m == 0 || (m & (m - 1))) return ctx->fail(SCZ_ERR_NOT_POW2, "acc_product: length %zu is not a power of two", m);
    if (m >= (1ull << 31)) return ctx->fail(SCZ_ERR_BAD_ARG, "acc_product: table too long");
    ProfScope ps(ctx, SCZ_K_ACC_PRODUCT);
    cudaStream_t st = ctx->stream;
    SCZ_CUDA(ctx, cudaMemcpyAsync(d_tree, d_x, m * 32, cudaMemcpyDeviceToDevice, st));
    size_t src = 0, dst = m, cnt = m / 2;
    while (cnt > (size_t)PL_THREADS) {
        k_tree_level<<<ceil_div_u32(cnt, PL_THREADS), PL_THREADS, 0, st>>>(d_tree, src, dst, (uint32_t)cnt);
        SCZ_LAUNCH_CHECK(ctx);
        src = dst;
        dst += cnt;
        cnt >>= 1;
    }
    k_tree_tail<<<1, PL_THREADS, 0, st>>>(d_tree, src, dst, (uint32_t)cnt, 2 * m - 1);
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}

// point-wise maps of dhyperplonk.rs:233-238, 251-256, 326-339 (modes: see scz_fr_pointwise_dev)
int32_t fr_pointwise(Ctx *c, int32_t mode, const void *d_a, const void *d_b, const void *d_k, void *d_out, size_t n) {
    if (!d_a || !d_b || !d_out || mode < 0 || mode > 3 || (mode == 2 && !d_k)) return c->fail(SCZ_ERR_BAD_ARG, "pointwise: bad argument");
    if (!n) return SCZ_OK;
    ProfScope ps(c, SCZ_K_POINTWISE);
    uint32_t g = ceil_div_u32(n, PL_THREADS);
    if (mode == 0) k_pointwise<0><<<g, PL_THREADS, 0, c->stream>>>(d_a, d_b, d_k, d_out, n);
    else if (mode == 1) k_pointwise<1><<<g, PL_THREADS, 0, c->stream>>>(d_a, d_b, d_k, d_out, n);
    else if (mode == 2) k_pointwise<2><<<g, PL_THREADS, 0, c->stream>>>(d_a, d_b, d_k, d_out, n);
    else {
        const uint32_t grid = ceil_div_u32(n, (size_t)DIV_THREADS * DIV_PER);
        DevTmp pre(c);
        InvTree<FrP> tree;
        SCZ_TRY(pre.alloc(n * sizeof(Fr)));
        SCZ_TRY(tree.alloc(c, (uint64_t)grid * DIV_THREADS));
        k_div_phase1<<<grid, DIV_THREADS, 0, c->stream>>>(d_b, n, pre.p, tree.values(), c->d_status);
        SCZ_LAUNCH_CHECK(c);
        SCZ_TRY(tree.run(c, grid * DIV_THREADS));
        k_div_phase2<<<grid, DIV_THREADS, 0, c->stream>>>(d_a, d_b, n, pre.p, tree.inverses(), d_out);
    }
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}

// even[i] = in[2i], odd[i] = in[2i+1]: v(x,0) / v(x,1) of the product tree (dhyperplonk.rs:349-359, dacc_product.rs:41-52)
__global__ void __launch_bounds__(PL_THREADS) k_deinterleave(const void *in, void *even, void *odd, size_t n_pairs) {
    size_t i = blockIdx.x * (size_t)PL_THREADS + threadIdx.x;
    if (i >= n_pairs) return;
    fp_store<FrP>(even, i, fp_load_rw<FrP>(in, 2 * i));
    fp_store<FrP>(odd, i, fp_load_rw<FrP>(in, 2 * i + 1));
}
int32_t fr_deinterleave(Ctx *c, const void *d_in, size_t n_pairs, void *d_even, void *d_odd) {
    if (!d_in || !d_even || !d_odd) return c->fail(SCZ_ERR_BAD_ARG, "deinterleave: null argument");
    if (!n_pairs) return SCZ_OK;
    ProfScope ps(c, SCZ_K_POINTWISE);
    k_deinterleave<<<ceil_div_u32(n_pairs, PL_THREADS), PL_THREADS, 0, c->stream>>>(d_in, d_even, d_odd, n_pairs);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_sumcheck_product_rounds_dev(scz_ctx *h, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                                        void *d_out_triples, void *d_last_fg) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (!d_f || !d_g || !d_last_fg || (len > 1 && (!d_challenge || !d_out_triples)))
        return h->c.fail(SCZ_ERR_BAD_ARG, "sumcheck: null argument");
    return sumcheck_product_rounds(&h->c, d_f, d_g, len, d_challenge, d_out_triples, d_last_fg);
}
int32_t scz_open_fold_dev(scz_ctx *h, const void *d_peval, size_t len, const void *d_point, void *d_q, void *d_value) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (!d_peval || !d_value || (len > 1 && (!d_point || !d_q))) return h->c.fail(SCZ_ERR_BAD_ARG, "open_fold: null argument");
    return open_fold_rounds(&h->c, d_peval, len, d_point, d_q, d_value);
}
int32_t scz_fix_variable_dev(scz_ctx *h, const void *d_evals, size_t len, const void *d_points, size_t npoints, void *d_out) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (!d_evals || !d_out || (npoints && !d_points)) return h->c.fail(SCZ_ERR_BAD_ARG, "fix_variable: null argument");
    return fix_variable_rounds(&h->c, d_evals, len, d_points, npoints, d_out);
}
int32_t scz_acc_product_dev(scz_ctx *h, const void *d_x, size_t m, void *d_tree) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (!d_x || !d_tree) return h->c.fail(SCZ_ERR_BAD_ARG, "acc_product: null argument");
    return acc_product_tree(&h->c, d_x, m, d_tree);
}
int32_t scz_fr_pointwise_dev(scz_ctx *h, int32_t mode, const void *d_a, const void *d_b, const void *d_k, void *d_out,
                             size_t n) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    return fr_pointwise(&h->c, mode, d_a, d_b, d_k, d_out, n);
}
int32_t scz_fr_deinterleave_dev(scz_ctx *h, const void *d_in, size_t n_pairs, void *d_even, void *d_odd) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    return fr_deinterleave(&h->c, d_in, n_pairs, d_even, d_odd);
}

}   // extern "C"
