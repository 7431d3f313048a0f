// MSM and d_msm over BLS12-381 G2.
//
// The reference's `d_msm` is generic over `G: CurveGroup` (dist-primitive/src/dmsm.rs:9-15) and BASELINE.json names
// "d_msm over G1/G2"; no caller of the reference instantiates it with G2 (dpoly_comm.rs:265, examples/msm.rs:66,89 use G1;
// G2 only appears in the SRS, dpoly_comm.rs:27,59-62).  This file completes the API: same semantics, same bit-exact
// results, a plain Pippenger that shares the G1 path's curve-independent front (digit recoding + counting sort,
// msm.cu: msm_sort_entries) and is NOT tuned like the G1 pipeline:
//   accumulate   one thread per bucket walks its run of the sorted entry stream (XYZZ += affine over Fq2)
//   windows      one CTA per (segment, window): 32 threads take contiguous bucket ranges (local sum S_t and weighted sum
//                T_t), thread 0 combines them: sum_j (j + 1) B_j = sum_t T_t + len * sum_t t * S_t
//   finish       one thread per segment: Horner over the window sums
// The leader closure of d_msm (dmsm.rs:31-38: unpack2, sum of the l secrets, replicate, pack) is the same rank-one map as
// over G1 (pss.h: u | p built at scz_pp_new): out_o = p_o * sum_j u_j * in_j, three small kernels of 255-bit
// double-and-add chains.
#include <string.h>

#include <algorithm>
#include <vector>

#include "g2.cuh"
#include "msm.h"
#include "msm_digits.cuh"
#include "net.h"
#include "pss.h"

namespace scz {

constexpr int G2_THREADS = 64;
constexpr int G2_WIN_THREADS = 32;   // two XYZZ values per thread in shared memory: 32 x 768 B
constexpr uint32_t G2_MAX_C = 12;   // at most 2^11 buckets per window: the window pass walks them with 64 threads

__global__ void __launch_bounds__(G2_THREADS) k_g2_accumulate(const MsmSeg *segs, int K, uint32_t buckets, const uint2 *sorted,
                                                               const uint32_t *counts, const uint32_t *cursor, void *bucket_out) {
    uint32_t b = blockIdx.x * G2_THREADS + threadIdx.x;
    if (b >= buckets) return;
    const uint32_t cnt = counts[b], end = cursor[b];
    G2X acc = G2X::inf();
    if (cnt) {
        int s = K == 1 ? 0 : seg_by_bucket(segs, K, b);
        const void *bases = segs[s].bases;
        for (uint32_t i = end - cnt; i < end; i++) {
            const uint2 e = sorted[i];
            G2Affine p = g2a_load(bases, e.x & 0x7fffffffu);
            g2x_add_affine(acc, p, (e.x >> 31) != 0);
        }
    }
    g2x_store(bucket_out, b, acc);
}

// window sum sum_j (j + 1) B_j of one (segment, window) = blockIdx.x
__global__ void __launch_bounds__(G2_WIN_THREADS) k_g2_windows(const MsmSeg *segs, int K, const void *bucket_in, void *wsum) {
    __shared__ uint32_t seg_of;
    if (threadIdx.x == 0) {
        int lo = 0, hi = K - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (segs[mid].window_base <= blockIdx.x) lo = mid;
            else hi = mid - 1;
        }
        seg_of = (uint32_t)lo;
    }
    __syncthreads();
    const MsmSeg sg = segs[seg_of];
    const uint32_t w = blockIdx.x - sg.window_base;
    const uint32_t nb = sg.nb, b0 = sg.bucket_base + w * nb;
    // contiguous ranges of `len` buckets per thread (len a power of two, at least 1)
    const uint32_t len = nb >= G2_WIN_THREADS ? nb / G2_WIN_THREADS : 1, nt = nb / len;
    __shared__ G2X S_sh[G2_WIN_THREADS], T_sh[G2_WIN_THREADS];
    if (threadIdx.x < nt) {
        G2X S = G2X::inf(), T = G2X::inf();   // S = sum B_j, T = sum (j - lo + 1) B_j over the range (running sums, top down)
        const uint32_t lo = threadIdx.x * len;
        for (uint32_t j = len; j-- > 0;) {
            S = g2x_add(S, g2x_load(bucket_in, b0 + lo + j));
            T = g2x_add(T, S);
        }
        S_sh[threadIdx.x] = S;
        T_sh[threadIdx.x] = T;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // total = sum_t T_t + len * sum_t t * S_t ; sum_t t * S_t by running sums over t = nt-1 .. 1
        G2X A = G2X::inf(), B = G2X::inf(), total = G2X::inf();
        for (uint32_t t = nt; t-- > 0;) {
            total = g2x_add(total, T_sh[t]);
            if (t) {
                A = g2x_add(A, S_sh[t]);
                B = g2x_add(B, A);
            }
        }
        for (uint32_t l = len; l > 1; l >>= 1) B = g2x_double(B);
        total = g2x_add(total, B);
        g2x_store(wsum, blockIdx.x, total);
    }
}

__global__ void k_g2_finish(const MsmSeg *segs, int K, const void *wsum, void *out_jac) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (uint32_t)K) return;
    const MsmSeg sg = segs[s];
    G2X acc = G2X::inf();
    if (sg.len)
        for (uint32_t w = sg.W; w-- > 0;) {
            for (uint32_t i = 0; i < sg.c; i++) acc = g2x_double(acc);
            acc = g2x_add(acc, g2x_load(wsum, sg.window_base + w));
        }
    g2j_store(sg.out ? sg.out : out_jac, sg.out ? 0 : s, g2x_to_jac(acc));
}

int32_t msm_g2_batched(Ctx *ctx, const void *const *d_bases, const void *const *d_scalars, const size_t *lens, size_t batch,
                       void *d_out) {
    if (batch == 0) return SCZ_OK;
    if (batch > (1u << 16)) return ctx->fail(SCZ_ERR_BAD_ARG, "msm_g2: batch too large");
    std::vector<MsmSeg> segs(batch);
    uint64_t points = 0, buckets = 0, windows = 0, entries = 0;
    for (size_t k = 0; k < batch; k++) {
        if (lens[k] && (!d_bases[k] || !d_scalars[k])) return ctx->fail(SCZ_ERR_BAD_ARG, "msm_g2: null segment %zu", k);
        MsmSeg &s = segs[k];
        memset(&s, 0, sizeof s);
        s.bases = d_bases[k];
        s.scalars = d_scalars[k];
        s.len = (uint32_t)lens[k];
        s.point_base = (uint32_t)points;
        s.pre = 0;
        s.c = std::min(ctx->msm_window_override ? ctx->msm_window_override : msm_pick_window(lens[k]), G2_MAX_C);
        s.Wd = s.W = msm_num_windows(s.c);
        s.nb = 1u << (s.c - 1);
        s.bucket_base = (uint32_t)buckets;
        s.window_base = (uint32_t)windows;
        s.out = nullptr;
        points += s.len;
        buckets += (uint64_t)s.W * s.nb;
        windows += s.W;
        entries += (uint64_t)s.len * s.Wd;
    }
    if (points >= (1ull << 31) || buckets >= (1ull << 31) || entries >= (1ull << 32))
        return ctx->fail(SCZ_ERR_BAD_ARG, "msm_g2: batch too large (%llu points)", (unsigned long long)points);
    cudaStream_t st = ctx->stream;
    DevTmp d_segs(ctx), d_counts(ctx), d_cursor(ctx), d_tiles(ctx), d_sorted(ctx), d_buckets(ctx), d_wsum(ctx);
    SCZ_TRY(d_segs.alloc(batch * sizeof(MsmSeg)));
    SCZ_TRY(d_counts.alloc(buckets * 4));
    SCZ_TRY(d_cursor.alloc(buckets * 4));
    SCZ_TRY(d_tiles.alloc((size_t)msm_scan_tiles(buckets) * 4 + 4));
    SCZ_TRY(d_sorted.alloc((entries ? entries : 1) * sizeof(uint2)));
    SCZ_TRY(d_buckets.alloc(buckets * sizeof(G2X)));
    SCZ_TRY(d_wsum.alloc(windows * sizeof(G2X)));
    SCZ_TRY(ctx->h2d_staged(d_segs.p, segs.data(), batch * sizeof(MsmSeg)));
    SCZ_CUDA(ctx, cudaMemsetAsync(d_counts.p, 0, buckets * 4, st));
    const MsmSeg *sp = d_segs.as<MsmSeg>();
    const int K = (int)batch;
    SCZ_TRY(msm_sort_entries(ctx, sp, K, (uint32_t)points, (uint32_t)buckets, d_counts.as<uint32_t>(), d_cursor.as<uint32_t>(),
                             d_tiles.as<uint32_t>(), d_sorted.as<uint2>()));
    k_g2_accumulate<<<ceil_div_u32(buckets, G2_THREADS), G2_THREADS, 0, st>>>(sp, K, (uint32_t)buckets, d_sorted.as<uint2>(),
                                                                              d_counts.as<uint32_t>(), d_cursor.as<uint32_t>(),
                                                                              d_buckets.p);
    SCZ_LAUNCH_CHECK(ctx);
    k_g2_windows<<<(uint32_t)windows, G2_WIN_THREADS, 0, st>>>(sp, K, d_buckets.p, d_wsum.p);
    SCZ_LAUNCH_CHECK(ctx);
    k_g2_finish<<<ceil_div_u32(batch, 32), 32, 0, st>>>(sp, K, d_wsum.p, d_out);
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}

// ---- leader closure of d_msm over G2: in / out party-major [party][k], n x batch Jacobian points (288 B)
__global__ void __launch_bounds__(G2_THREADS) k_g2_closure_scale_in(const void *UP, uint32_t n, uint32_t batch, const void *in, void *part) {
    uint32_t t = blockIdx.x * G2_THREADS + threadIdx.x;   // t = j * batch + k
    if (t >= n * batch) return;
    uint32_t j = t / batch;
    Fr u = fp_to_canon(fp_load<FrP>(UP, j));
    g2x_store(part, t, g2x_mul_bits(g2x_from_jac(g2j_load(in, t)), u.l));
}
__global__ void __launch_bounds__(G2_THREADS) k_g2_closure_sum(uint32_t n, uint32_t batch, const void *part, void *S) {
    uint32_t k = blockIdx.x * G2_THREADS + threadIdx.x;
    if (k >= batch) return;
    G2X acc = G2X::inf();
    for (uint32_t j = 0; j < n; j++) acc = g2x_add(acc, g2x_load(part, (size_t)j * batch + k));
    g2x_store(S, k, acc);
}
__global__ void __launch_bounds__(G2_THREADS) k_g2_closure_scale_out(const void *UP, uint32_t n, uint32_t batch, const void *S, void *out) {
    uint32_t t = blockIdx.x * G2_THREADS + threadIdx.x;   // t = o * batch + k
    if (t >= n * batch) return;
    uint32_t o = t / batch, k = t - o * batch;
    Fr p = fp_to_canon(fp_load<FrP>(UP, (size_t)n + o));
    g2j_store(out, t, g2x_to_jac(g2x_mul_bits(g2x_load(S, k), p.l)));
}
int32_t d_msm_g2_leader(Ctx *ctx, const scz_pp *pp, const void *d_recv, size_t batch, void *d_send) {
    ProfScope ps(ctx, SCZ_K_PSS);
    const uint32_t n = (uint32_t)pp->n, B = (uint32_t)batch;
    DevTmp part(ctx), S(ctx);
    SCZ_TRY(part.alloc((size_t)n * B * sizeof(G2X)));
    SCZ_TRY(S.alloc((size_t)B * sizeof(G2X)));
    cudaStream_t st = ctx->stream;
    k_g2_closure_scale_in<<<ceil_div_u32((size_t)n * B, G2_THREADS), G2_THREADS, 0, st>>>(pp->d_dmsm, n, B, d_recv, part.p);
    SCZ_LAUNCH_CHECK(ctx);
    k_g2_closure_sum<<<ceil_div_u32(B, G2_THREADS), G2_THREADS, 0, st>>>(n, B, part.p, S.p);
    SCZ_LAUNCH_CHECK(ctx);
    k_g2_closure_scale_out<<<ceil_div_u32((size_t)n * B, G2_THREADS), G2_THREADS, 0, st>>>(pp->d_dmsm, n, B, S.p, d_send);
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}

// d_msm over G2 (dmsm.rs:9-43): local MSMs, gather, leader closure, scatter.  Wire size of Vec<G2>: 8 + 96 per point
// (ark-bls12-381's compressed G2 encoding)
int32_t d_msm_g2_dev(Ctx *ctx, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars, const size_t *lens,
                     size_t batch, void *d_out) {
    if (!pp) return ctx->fail(SCZ_ERR_BAD_ARG, "d_msm_g2: null pp");
    if (batch == 0) return SCZ_OK;
    Net *net = ctx->net;
    const size_t N = net->n_parties, PT = sizeof(G2Jac);
    if (N != pp->n) return ctx->fail(SCZ_ERR_BAD_ARG, "d_msm_g2: %zu parties but pp.n = %zu", N, pp->n);
    DevTmp c_shares(ctx), recv(ctx), send(ctx);
    SCZ_TRY(c_shares.alloc(batch * PT));
    SCZ_TRY(msm_g2_batched(ctx, d_bases, d_scalars, lens, batch, c_shares.p));        // :19-24
    const size_t wire = 8 + 96 * batch;
    if (net->is_leader()) {
        SCZ_TRY(recv.alloc(N * batch * PT));
        SCZ_TRY(send.alloc(N * batch * PT));
    }
    SCZ_TRY(net->gather(ctx, c_shares.p, recv.p, batch * PT, wire));                  // :29
    if (net->is_leader()) SCZ_TRY(d_msm_g2_leader(ctx, pp, recv.p, batch, send.p));  // :31-38
    return net->scatter(ctx, send.p, d_out, batch * PT, wire);                        // :40
}

__global__ void k_g2_apply_inf_mask(void *bases, const uint8_t *mask, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n || !mask[i]) return;
    G2Affine z;
    z.x = Fq2::zero();
    z.y = Fq2::zero();
    g2a_store(bases, i, z);
}

__global__ void __launch_bounds__(G2_THREADS) k_g2_vec_op(int op, const void *a, const void *b, void *out, size_t n) {
    size_t i = blockIdx.x * (size_t)G2_THREADS + threadIdx.x;
    if (i >= n) return;
    G2X x = g2x_from_jac(g2j_load(a, i)), r;
    if (op == 0) r = g2x_add(x, g2x_from_jac(g2j_load(b, i)));
    else if (op == 1) r = g2x_double(x);
    else {
        Fr k = fp_to_canon(fp_load<FrP>(b, i));
        r = g2x_mul_bits(x, k.l);
    }
    g2j_store(out, i, g2x_to_jac(r));
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_msm_g2_batched_dev(scz_ctx *h, const void *const *d_bases, const void *const *d_scalars, const size_t *lens,
                               size_t batch, void *d_out) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (batch && (!d_bases || !d_scalars || !lens || !d_out)) return h->c.fail(SCZ_ERR_BAD_ARG, "msm_g2: null argument");
    return msm_g2_batched(&h->c, d_bases, d_scalars, lens, batch, d_out);
}

int32_t scz_msm_g2(scz_ctx *h, const void *bases, const uint8_t *inf_mask, size_t bases_len, const void *scalars,
                   size_t scalars_len, void *out_jac) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (!out_jac) return c->fail(SCZ_ERR_BAD_ARG, "msm_g2: null output");
    if (bases_len != scalars_len)   // ark-ec returns Err(min len); the reference unwrap()s it (dmsm.rs:23)
        return c->fail(SCZ_ERR_LEN_MISMATCH, "msm_g2: %zu bases vs %zu scalars", bases_len, scalars_len);
    size_t n = bases_len;
    if (n && (!bases || !scalars)) return c->fail(SCZ_ERR_BAD_ARG, "msm_g2: null input");
    DevTmp d_b(c), d_s(c), d_m(c), d_o(c);
    SCZ_TRY(d_b.alloc(n * SCZ_G2_AFFINE_BYTES));
    SCZ_TRY(d_s.alloc(n * SCZ_FR_BYTES));
    SCZ_TRY(d_o.alloc(SCZ_G2_JAC_BYTES));
    if (n) {
        SCZ_CUDA(c, cudaMemcpyAsync(d_b.p, bases, n * SCZ_G2_AFFINE_BYTES, cudaMemcpyHostToDevice, c->stream));
        SCZ_CUDA(c, cudaMemcpyAsync(d_s.p, scalars, n * SCZ_FR_BYTES, cudaMemcpyHostToDevice, c->stream));
        if (inf_mask) {
            SCZ_TRY(d_m.alloc(n));
            SCZ_CUDA(c, cudaMemcpyAsync(d_m.p, inf_mask, n, cudaMemcpyHostToDevice, c->stream));
            k_g2_apply_inf_mask<<<ceil_div_u32(n, 256), 256, 0, c->stream>>>(d_b.p, d_m.as<uint8_t>(), n);
            SCZ_LAUNCH_CHECK(c);
        }
    }
    const void *bp = d_b.p, *sp = d_s.p;
    SCZ_TRY(msm_g2_batched(c, &bp, &sp, &n, 1, d_o.p));
    SCZ_CUDA(c, cudaMemcpyAsync(out_jac, d_o.p, SCZ_G2_JAC_BYTES, cudaMemcpyDeviceToHost, c->stream));
    SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    return SCZ_OK;
}

int32_t scz_d_msm_g2_dev(scz_ctx *h, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars,
                         const size_t *lens, size_t batch, void *d_out) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (batch && (!d_bases || !d_scalars || !lens || !d_out)) return h->c.fail(SCZ_ERR_BAD_ARG, "d_msm_g2: null argument");
    return d_msm_g2_dev(&h->c, pp, d_bases, d_scalars, lens, batch, d_out);
}

int32_t scz_d_msm_g2_leader_dev(scz_ctx *h, const scz_pp *pp, const void *d_gathered, size_t batch, void *d_to_scatter) {
    scz::DeviceGuard dg__(h);
    if (!h || !pp) return SCZ_ERR_BAD_ARG;
    if (batch && (!d_gathered || !d_to_scatter)) return h->c.fail(SCZ_ERR_BAD_ARG, "d_msm_g2_leader: null argument");
    if (!batch) return SCZ_OK;
    return d_msm_g2_leader(&h->c, pp, d_gathered, batch, d_to_scatter);
}

// G2 unit operations for the parity tests: op 0: out = a + b (Jacobian), 1: out = 2 a, 2: out = k * a (d_b: Fr, Montgomery)
int32_t scz_g2_vec_op_dev(scz_ctx *h, int32_t op, const void *d_a_jac, const void *d_b, void *d_out_jac, size_t n) {
    scz::DeviceGuard dg__(h);
    if (!h || op < 0 || op > 2 || (n && (!d_a_jac || !d_out_jac || (op != 1 && !d_b)))) return SCZ_ERR_BAD_ARG;
    if (!n) return SCZ_OK;
    k_g2_vec_op<<<ceil_div_u32(n, G2_THREADS), G2_THREADS, 0, h->c.stream>>>(op, d_a_jac, d_b, d_out_jac, n);
    h->c.launches++;
    return cudaGetLastError() == cudaSuccess ? SCZ_OK : SCZ_ERR_CUDA;
}

}   // extern "C"
