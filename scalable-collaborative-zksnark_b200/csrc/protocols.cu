// The dist-primitive protocols on top of the kernels: every function here is the device-side restatement
// of one reference function, local loops on this party's GPU and the reference's own leader rounds
// (gather -> closure on the leader -> scatter, dist-primitive/src/utils/serializing_net.rs:128-141) through
// ctx->net.  Payloads stay on the device in device layout; the byte counters use the reference's
// ark-serialize compressed sizes (Fr 32 B, G1 48 B, Vec header 8 B) so get_comm() stays comparable.
//
//   pss2ss                 dist-primitive/src/unpack.rs:72-97
//   degree_reduce          dist-primitive/src/degree_reduce.rs:29-41
//   sumcheck_product       dist-primitive/src/dsumcheck.rs:28-90
//   c_sumcheck_product     dist-primitive/src/dsumcheck.rs:148-285
//   d_sumcheck_product     dist-primitive/src/dsumcheck.rs:359-512
//   commit / d_local_commit, c_commit, d_commit     dist-primitive/src/dpoly_comm.rs:237-243, 269-275, 244-267, 276-297
//   open / d_local_open, c_open, d_open             dist-primitive/src/dpoly_comm.rs:299-325, 327-353, 401-464, 355-398
//   d_acc_product          dist-primitive/src/dacc_product.rs:365-414
#include <vector>

#include "deferred.h"
#include "g1.cuh"
#include "msm.h"
#include "net.h"
#include "pss.h"
#include "srs.h"

namespace scz {

int32_t sumcheck_product_rounds(Ctx *ctx, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                                void *d_out, void *d_last);
int32_t open_fold_rounds(Ctx *ctx, const void *d_peval, size_t len, const void *d_point, void *d_q, void *d_value);
int32_t sumcheck_rounds(Ctx *ctx, const void *d_f, size_t len, const void *d_challenge, void *d_out, void *d_last);
int32_t acc_product_tree(Ctx *ctx, const void *d_x, size_t m, void *d_tree);
int32_t d_msm_defer(Ctx *ctx, Deferred &D, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars,
                    const size_t *lens, size_t batch, void *d_out, const uint32_t *pre_c = nullptr);

static inline size_t log2_exact(size_t v, bool *ok) {
    size_t l = 0;
    while (((size_t)1 << l) < v) l++;
    *ok = v != 0 && ((size_t)1 << l) == v;
    return l;
}

struct Fr3 {
    Fr a, b, c;
};

// ---- tiny leader-side kernels -------------------------------------------------------------------
// out = (0, f*g, 0)            dsumcheck.rs:87 / :282
__global__ void k_final_triple(const void *last_fg, void *out) {
    if (threadIdx.x) return;
    Fr f = fp_load_rw<FrP>(last_fg, 0), g = fp_load_rw<FrP>(last_fg, 1);
    fp_store<FrP>(out, 0, Fr::zero());
    fp_store<FrP>(out, 1, fp_mul(f, g));
    fp_store<FrP>(out, 2, Fr::zero());
}
// out = (g_last, f_last, 0)    dsumcheck.rs:433
__global__ void k_last_triple(const void *last_fg, void *out) {
    if (threadIdx.x) return;
    Fr f = fp_load_rw<FrP>(last_fg, 0), g = fp_load_rw<FrP>(last_fg, 1);
    fp_store<FrP>(out, 0, g);
    fp_store<FrP>(out, 1, f);
    fp_store<FrP>(out, 2, Fr::zero());
}
// leader of d_sumcheck_product (dsumcheck.rs:440-449): recv is [party][n+1] triples.
// out[i] = sum over parties of round i; lf[j] = recv[j][n].b ; lg[j] = recv[j][n].a
__global__ void k_dsum_leader(const void *recv, uint32_t N, uint32_t n, void *out, void *lf, void *lg) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 3 * n) {
        uint32_t i = t / 3, comp = t % 3;
        Fr acc = Fr::zero();
        for (uint32_t j = 0; j < N; j++) acc = fp_add(acc, fp_load_rw<FrP>(recv, ((size_t)j * (n + 1) + i) * 3 + comp));
        fp_store<FrP>(out, (size_t)i * 3 + comp, acc);
    } else if (t < 3 * n + N) {
        uint32_t j = t - 3 * n;
        fp_store<FrP>(lg, j, fp_load_rw<FrP>(recv, ((size_t)j * (n + 1) + n) * 3));
        fp_store<FrP>(lf, j, fp_load_rw<FrP>(recv, ((size_t)j * (n + 1) + n) * 3 + 1));
    }
}
// lz[j] = first Fr of party j's payload
__global__ void k_pick_fr(const void *in, size_t party_stride_bytes, uint32_t N, void *out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    fp_store<FrP>(out, j, fp_load_rw<FrP>(reinterpret_cast<const char *>(in) + (size_t)j * party_stride_bytes, 0));
}

// ---- pss2ss: unpack.rs:72-97 ----------------------------------------------------------------------
// gather one share -> leader: unpack -> pack_single each secret -> transpose -> scatter Vec<F> of length l
int32_t pss2ss_dev(Ctx *ctx, const scz_pp *pp, const void *d_share, void *d_out) {
    Net *net = ctx->net;
    const size_t N = net->n_parties, l = pp->l;
    if (N != pp->n) return ctx->fail(SCZ_ERR_BAD_ARG, "pss2ss: %zu parties but pp.n = %zu", N, pp->n);
    DevTmp recv(ctx), sec(ctx), send(ctx);
    if (net->is_leader()) {
        SCZ_TRY(recv.alloc(N * 32));
        SCZ_TRY(sec.alloc(l * 32));
        SCZ_TRY(send.alloc(N * l * 32));
    }
    SCZ_TRY(net->gather(ctx, d_share, recv.p, 32, 32));
    if (net->is_leader()) {
        ProfScope ps(ctx, SCZ_K_PSS);
        SCZ_TRY(pss_apply(ctx, pp, PSS_UNPACK, 0, recv.p, N, N, 1, 1, sec.p, l, 1));
        // secret i -> shares PS[j] * sec[i], stored party-major [j][i]
        SCZ_TRY(pss_apply(ctx, pp, PSS_PACK_SINGLE, 0, sec.p, 1, 1, 1, l, send.p, 1, l));
    }
    SCZ_TRY(net->scatter(ctx, send.p, d_out, l * 32, 8 + 32 * l));
    return SCZ_OK;
}

// k pss2ss calls on k shares in ONE leader round (c_sumcheck_product does two back to back, dsumcheck.rs:224-225):
// d_out is [k][l].  Byte counters as for k separate rounds.
int32_t pss2ss_many_dev(Ctx *ctx, const scz_pp *pp, const void *d_shares, size_t k, void *d_out) {
    Net *net = ctx->net;
    const size_t N = net->n_parties, l = pp->l;
    if (N != pp->n) return ctx->fail(SCZ_ERR_BAD_ARG, "pss2ss: %zu parties but pp.n = %zu", N, pp->n);
    DevTmp recv(ctx), sec(ctx), send(ctx);
    if (net->is_leader()) {
        SCZ_TRY(recv.alloc(N * k * 32));
        SCZ_TRY(sec.alloc(k * l * 32));
        SCZ_TRY(send.alloc(N * k * l * 32));
    }
    SCZ_TRY(net->gather(ctx, d_shares, recv.p, k * 32, k * 32));
    if (net->is_leader()) {
        ProfScope ps(ctx, SCZ_K_PSS);
        SCZ_TRY(pss_apply(ctx, pp, PSS_UNPACK, 0, recv.p, N, 1, k, k, sec.p, l, 1));
        SCZ_TRY(pss_apply(ctx, pp, PSS_PACK_SINGLE, 0, sec.p, 1, 1, 1, k * l, send.p, 1, k * l));
    }
    SCZ_TRY(net->scatter(ctx, send.p, d_out, k * l * 32, k * (8 + 32 * l)));
    return SCZ_OK;
}

// ---- degree_reduce: degree_reduce.rs:29-41 ----------------------------------------------------------
int32_t degree_reduce_dev(Ctx *ctx, const scz_pp *pp, const void *d_share, void *d_out) {
    Net *net = ctx->net;
    const size_t N = net->n_parties, l = pp->l;
    if (N != pp->n) return ctx->fail(SCZ_ERR_BAD_ARG, "degree_reduce: %zu parties but pp.n = %zu", N, pp->n);
    DevTmp recv(ctx), sec(ctx), send(ctx);
    if (net->is_leader()) {
        SCZ_TRY(recv.alloc(N * 32));
        SCZ_TRY(sec.alloc(l * 32));
        SCZ_TRY(send.alloc(N * 32));
    }
    SCZ_TRY(net->gather(ctx, d_share, recv.p, 32, 32));
    if (net->is_leader()) {
        ProfScope ps(ctx, SCZ_K_PSS);
        SCZ_TRY(pss_apply(ctx, pp, PSS_UNPACK2, 0, recv.p, N, N, 1, 1, sec.p, l, 1));
        SCZ_TRY(pss_apply(ctx, pp, PSS_PACK, 0, sec.p, l, l, 1, 1, send.p, N, 1));
    }
    SCZ_TRY(net->scatter(ctx, send.p, d_out, 32, 32));
    return SCZ_OK;
}

// ---- sumcheck_product: dsumcheck.rs:28-90 -> n + 1 triples ----------------------------------------------
int32_t sumcheck_product_dev(Ctx *ctx, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                             void *d_out) {
    bool ok;
    size_t n = log2_exact(len, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "sumcheck_product: length %zu is not a power of two", len);
    DevTmp last(ctx);
    SCZ_TRY(last.alloc(64));
    SCZ_TRY(sumcheck_product_rounds(ctx, d_f, d_g, len, d_challenge, d_out, last.p));
    k_final_triple<<<1, 32, 0, ctx->stream>>>(last.p, (char *)d_out + n * 96);
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}

// ---- c_sumcheck_product: dsumcheck.rs:148-285 -> n + log2(l) + 1 triples ---------------------------------
int32_t c_sumcheck_product_dev(Ctx *ctx, const scz_pp *pp, const void *d_f, const void *d_g, size_t len,
                               const void *d_challenge, void *d_out) {
    bool ok;
    size_t n = log2_exact(len, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "c_sumcheck_product: length %zu is not a power of two", len);
    size_t l = pp->l, ll = log2_exact(l, &ok);
    DevTmp last(ctx), fg2(ctx), last2(ctx);
    SCZ_TRY(last.alloc(64));
    SCZ_TRY(fg2.alloc(2 * l * 32));
    SCZ_TRY(last2.alloc(64));
    SCZ_TRY(sumcheck_product_rounds(ctx, d_f, d_g, len, d_challenge, d_out, last.p));   // Phase 1 :167-219
    SCZ_TRY(pss2ss_many_dev(ctx, pp, last.p, 2, fg2.p));                                // :224-225, one round for both
    // Phase 2 :227-279 -- indexes challenge[i] with i from 0 again (:230), replicated as is
    SCZ_TRY(sumcheck_product_rounds(ctx, fg2.p, (char *)fg2.p + l * 32, l, d_challenge, (char *)d_out + n * 96, last2.p));
    k_final_triple<<<1, 32, 0, ctx->stream>>>(last2.p, (char *)d_out + (n + ll) * 96);   // :282
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}

// ---- d_sumcheck_product: dsumcheck.rs:359-512.  Leader: n + log2(N) triples, others: none (:507-509) -----------
int32_t d_sumcheck_product_dev(Ctx *ctx, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                               void *d_out, size_t *count) {
    bool ok;
    size_t n = log2_exact(len, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "d_sumcheck_product: length %zu is not a power of two", len);
    Net *net = ctx->net;
    const size_t N = net->n_parties;
    size_t s = log2_exact(N, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "d_sumcheck_product: %zu parties is not a power of two", N);
    DevTmp local(ctx), last(ctx), recv(ctx), lf(ctx), lg(ctx), last2(ctx);
    SCZ_TRY(local.alloc((n + 1) * 96));
    SCZ_TRY(last.alloc(64));
    SCZ_TRY(sumcheck_product_rounds(ctx, d_f, d_g, len, d_challenge, local.p, last.p));   // :377-429
    k_last_triple<<<1, 32, 0, ctx->stream>>>(last.p, (char *)local.p + n * 96);           // :433
    SCZ_LAUNCH_CHECK(ctx);
    if (net->is_leader()) {
        SCZ_TRY(recv.alloc(N * (n + 1) * 96));
        SCZ_TRY(lf.alloc(N * 32));
        SCZ_TRY(lg.alloc(N * 32));
        SCZ_TRY(last2.alloc(64));
    }
    SCZ_TRY(net->gather(ctx, local.p, recv.p, (n + 1) * 96, 8 + 96 * (n + 1)));            // :434
    if (!net->is_leader()) {
        if (count) *count = 0;
        return SCZ_OK;
    }
    uint32_t work = (uint32_t)(3 * n + N);
    k_dsum_leader<<<ceil_div_u32(work, 64), 64, 0, ctx->stream>>>(recv.p, (uint32_t)N, (uint32_t)n, d_out, lf.p, lg.p);
    SCZ_LAUNCH_CHECK(ctx);
    SCZ_TRY(sumcheck_product_rounds(ctx, lf.p, lg.p, N, (const char *)d_challenge + n * 32, (char *)d_out + n * 96,
                                    last2.p));                                            // :452-504
    if (count) *count = n + s;
    return SCZ_OK;
}

// ---- single-MLE sumchecks: dsumcheck.rs:6-26, 92-146, 287-357 (round message = (sum lo, sum hi), 64 B) --------------
// out = (0, v)
__global__ void k_final_pair(const void *last, void *out) {
    if (threadIdx.x) return;
    fp_store<FrP>(out, 0, Fr::zero());
    fp_store<FrP>(out, 1, fp_load_rw<FrP>(last, 0));
}
// leader of d_sumcheck (:321-333): recv is [party][n+1] pairs.  out[i] = sum over parties of round i; lf[j] = recv[j][n].b
__global__ void k_dsum1_leader(const void *recv, uint32_t N, uint32_t n, void *out, void *lf) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 2 * n) {
        uint32_t i = t / 2, comp = t % 2;
        Fr acc = Fr::zero();
        for (uint32_t j = 0; j < N; j++) acc = fp_add(acc, fp_load_rw<FrP>(recv, ((size_t)j * (n + 1) + i) * 2 + comp));
        fp_store<FrP>(out, (size_t)i * 2 + comp, acc);
    } else if (t < 2 * n + N) {
        uint32_t j = t - 2 * n;
        fp_store<FrP>(lf, j, fp_load_rw<FrP>(recv, ((size_t)j * (n + 1) + n) * 2 + 1));
    }
}
// sumcheck (:6-26) -> n + 1 pairs, the last is (0, f(challenge))
int32_t sumcheck_dev(Ctx *ctx, const void *d_f, size_t len, const void *d_challenge, void *d_out) {
    bool ok;
    size_t n = log2_exact(len, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "sumcheck: length %zu is not a power of two", len);
    DevTmp last(ctx);
    SCZ_TRY(last.alloc(32));
    SCZ_TRY(sumcheck_rounds(ctx, d_f, len, d_challenge, d_out, last.p));
    k_final_pair<<<1, 32, 0, ctx->stream>>>(last.p, (char *)d_out + n * 64);
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}
// c_sumcheck (:92-146) -> n + log2(l) + 1 pairs; phase 2 re-uses challenge[0..log2 l) (:129) like the product variant
int32_t c_sumcheck_dev(Ctx *ctx, const scz_pp *pp, const void *d_f, size_t len, const void *d_challenge, void *d_out) {
    bool ok;
    size_t n = log2_exact(len, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "c_sumcheck: length %zu is not a power of two", len);
    size_t l = pp->l, ll = log2_exact(l, &ok);
    DevTmp last(ctx), f2(ctx), last2(ctx);
    SCZ_TRY(last.alloc(32));
    SCZ_TRY(f2.alloc(l * 32));
    SCZ_TRY(last2.alloc(32));
    SCZ_TRY(sumcheck_rounds(ctx, d_f, len, d_challenge, d_out, last.p));                  // Phase 1 :105-121
    SCZ_TRY(pss2ss_dev(ctx, pp, last.p, f2.p));                                           // :124
    SCZ_TRY(sumcheck_rounds(ctx, f2.p, l, d_challenge, (char *)d_out + n * 64, last2.p)); // Phase 2 :127-141
    k_final_pair<<<1, 32, 0, ctx->stream>>>(last2.p, (char *)d_out + (n + ll) * 64);      // :143
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}
// d_sumcheck (:287-357).  Leader: n + log2(N) pairs, others: none (:351-353)
int32_t d_sumcheck_dev(Ctx *ctx, const void *d_f, size_t len, const void *d_challenge, void *d_out, size_t *count) {
    bool ok;
    size_t n = log2_exact(len, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "d_sumcheck: length %zu is not a power of two", len);
    Net *net = ctx->net;
    const size_t N = net->n_parties;
    size_t s = log2_exact(N, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "d_sumcheck: %zu parties is not a power of two", N);
    DevTmp local(ctx), last(ctx), recv(ctx), lf(ctx), last2(ctx);
    SCZ_TRY(local.alloc((n + 1) * 64));
    SCZ_TRY(last.alloc(32));
    SCZ_TRY(sumcheck_rounds(ctx, d_f, len, d_challenge, local.p, last.p));                // :301-316
    k_final_pair<<<1, 32, 0, ctx->stream>>>(last.p, (char *)local.p + n * 64);            // :318
    SCZ_LAUNCH_CHECK(ctx);
    if (net->is_leader()) {
        SCZ_TRY(recv.alloc(N * (n + 1) * 64));
        SCZ_TRY(lf.alloc(N * 32));
        SCZ_TRY(last2.alloc(32));
    }
    SCZ_TRY(net->gather(ctx, local.p, recv.p, (n + 1) * 64, 8 + 64 * (n + 1)));            // :320-322
    if (!net->is_leader()) {
        if (count) *count = 0;
        return SCZ_OK;
    }
    uint32_t work = (uint32_t)(2 * n + N);
    k_dsum1_leader<<<ceil_div_u32(work, 64), 64, 0, ctx->stream>>>(recv.p, (uint32_t)N, (uint32_t)n, d_out, lf.p);
    SCZ_LAUNCH_CHECK(ctx);
    SCZ_TRY(sumcheck_rounds(ctx, lf.p, N, (const char *)d_challenge + n * 32, (char *)d_out + n * 64, last2.p));   // :334-347
    if (count) *count = n + s;
    return SCZ_OK;
}

// ---- d_acc_product: dacc_product.rs:365-414 ---------------------------------------------------------------
int32_t d_acc_product_dev(Ctx *ctx, const void *d_x, size_t m, void *d_subtree, void *d_leader_tree) {
    Net *net = ctx->net;
    const size_t N = net->n_parties;
    SCZ_TRY(acc_product_tree(ctx, d_x, m, d_subtree));                                    // :374-381
    DevTmp recv(ctx);
    if (net->is_leader()) SCZ_TRY(recv.alloc(N * 32));
    // the entry sent is subtree[2m-1], already forced to zero (:381, :390) -- replicated, not "fixed"
    SCZ_TRY(net->gather(ctx, (const char *)d_subtree + (2 * m - 1) * 32, recv.p, 32, 32));
    if (net->is_leader()) SCZ_TRY(acc_product_tree(ctx, recv.p, N, d_leader_tree));        // :392-404
    return SCZ_OK;
}

}   // namespace scz

// ---- SRS: the G1 side of PolynomialCommitment (dpoly_comm.rs:30-34) ------------------------------------------
namespace scz {

// bases of `level_of`'s level; when the level has a fixed-base table (srs.cu) *out is the table and *pre_c its window
static int32_t srs_level_for(Ctx *ctx, const scz_srs *srs, size_t need_len, size_t level_of, const char *who,
                             const void **out, uint32_t *pre_c) {
    bool ok;
    size_t level = log2_exact(level_of, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "%s: length %zu is not a power of two", who, level_of);   // dpoly_comm.rs:240,255
    if (level >= srs->level.size()) return ctx->fail(SCZ_ERR_LEVEL_OOB, "%s: level %zu >= %zu", who, level, srs->level.size());
    if (srs->len[level] < need_len)
        return ctx->fail(SCZ_ERR_LEN_MISMATCH, "%s: level %zu holds %zu bases, %zu needed", who, level, srs->len[level], need_len);
    // a table is laid out [window][point] for the level's full length: usable when the MSM covers the whole level
    bool pre = level < srs->table.size() && srs->table[level] && srs->len[level] == need_len && !ctx->msm_no_precompute;
    *out = pre ? srs->table[level] : srs->level[level];
    *pre_c = pre ? srs->table_c[level] : 0;
    return SCZ_OK;
}

// Every function of the commit / open family comes as `*_defer` (queues its MSMs on `D`, registers its leader round
// as a continuation; see deferred.h) and as `*_dev` = defer + run for stand-alone calls.

// commit / d_local_commit: dpoly_comm.rs:237-243, 269-275
int32_t commit_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const void *d_peval, size_t len, void *d_out) {
    const void *b = nullptr;
    uint32_t pre = 0;
    SCZ_TRY(srs_level_for(ctx, srs, len, len, "commit", &b, &pre));
    return D.add_msm(b, d_peval, len, d_out, pre);
}

// c_commit: dpoly_comm.rs:244-267
int32_t c_commit_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const scz_pp *pp, const void *const *d_pevals,
                       const size_t *lens, size_t batch, void *d_out) {
    std::vector<const void *> bases(batch);
    std::vector<uint32_t> pre(batch);
    for (size_t k = 0; k < batch; k++)
        SCZ_TRY(srs_level_for(ctx, srs, lens[k], lens[k] * pp->l, "c_commit", &bases[k], &pre[k]));
    return d_msm_defer(ctx, D, pp, bases.data(), d_pevals, lens, batch, d_out, pre.data());
}

// d_commit: dpoly_comm.rs:276-297 -- every party ends with the sum of the N local commitments
int32_t d_commit_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const void *d_peval, size_t len, void *d_out) {
    Net *net = ctx->net;
    const size_t N = net->n_parties, PT = SCZ_G1_JAC_BYTES;
    DevTmp *loc = nullptr;
    SCZ_TRY(D.tmp(PT, &loc));
    SCZ_TRY(commit_defer(ctx, D, srs, d_peval, len, loc->p));
    Deferred *Dp = &D;
    D.then([=]() -> int32_t {
        DevTmp *recv = nullptr, *send = nullptr;
        if (net->is_leader()) {
            SCZ_TRY(Dp->tmp(N * PT, &recv));
            SCZ_TRY(Dp->tmp(N * PT, &send));
        }
        Dp->gather(loc->p, recv ? recv->p : nullptr, PT, 48);
        if (net->is_leader()) Dp->add_colsum(recv->p, PT, 0, (uint32_t)N, 1, send->p, (uint32_t)N);   // sum, N copies :290-292
        Dp->scatter(send ? send->p : nullptr, d_out, PT, 48);
        return SCZ_OK;
    });
    return SCZ_OK;
}

// open / d_local_open: dpoly_comm.rs:299-325, 327-353.  The reference commits q_i inside the fold loop; here all folds
// run first (now) and the n MSMs are queued (same values).
int32_t open_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point,
                   void *d_value, void *d_proofs) {
    bool ok;
    size_t n = log2_exact(len, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "open: length %zu is not a power of two", len);
    DevTmp *q = nullptr;
    SCZ_TRY(D.tmp((len > 1 ? len - 1 : 1) * 32, &q));
    SCZ_TRY(open_fold_rounds(ctx, d_peval, len, d_point, q->p, d_value));
    size_t off = 0;
    for (size_t i = 0; i < n; i++) {
        size_t h = len >> (i + 1);
        const void *b = nullptr;
        uint32_t pre = 0;
        SCZ_TRY(srs_level_for(ctx, srs, h, h, "open", &b, &pre));
        SCZ_TRY(D.add_msm(b, (const char *)q->p + off * 32, h, (char *)d_proofs + i * SCZ_G1_JAC_BYTES, pre));
        off += h;
    }
    return SCZ_OK;
}

// c_open: dpoly_comm.rs:401-464 -> value + n + log2(l) proofs
int32_t c_open_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const scz_pp *pp, const void *d_peval, size_t len,
                     const void *d_point, void *d_value, void *d_proofs) {
    bool ok;
    size_t n = log2_exact(len, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "c_open: length %zu is not a power of two", len);
    const size_t l = pp->l, ll = log2_exact(l, &ok), PT = SCZ_G1_JAC_BYTES;
    DevTmp *q = nullptr, *last = nullptr, *r2 = nullptr, *q2 = nullptr;
    SCZ_TRY(D.tmp((len > 1 ? len - 1 : 1) * 32, &q));
    SCZ_TRY(D.tmp(32, &last));
    SCZ_TRY(open_fold_rounds(ctx, d_peval, len, d_point, q->p, last->p));                   // Phase 1 :418-432
    if (n) {
        std::vector<const void *> scal(n);
        std::vector<size_t> lens(n);
        size_t off = 0;
        for (size_t i = 0; i < n; i++) {
            size_t h = len >> (i + 1);
            scal[i] = (const char *)q->p + off * 32;
            lens[i] = h;
            off += h;
        }
        SCZ_TRY(c_commit_defer(ctx, D, srs, pp, scal.data(), lens.data(), n, d_proofs));    // ONE batched c_commit :436
    }
    SCZ_TRY(D.tmp(l * 32, &r2));
    SCZ_TRY(pss2ss_dev(ctx, pp, last->p, r2->p));                                           // :439
    SCZ_TRY(D.tmp((l > 1 ? l - 1 : 1) * 32, &q2));
    SCZ_TRY(open_fold_rounds(ctx, r2->p, l, d_point, q2->p, d_value));                      // Phase 2 :442-459, point[i] from 0
    size_t off = 0;
    for (size_t i = 0; i < ll; i++) {
        size_t h = l >> (i + 1);
        const void *b = nullptr;
        uint32_t pre = 0;
        SCZ_TRY(srs_level_for(ctx, srs, h, h * l, "c_open", &b, &pre));                     // local G1::msm :457
        SCZ_TRY(D.add_msm(b, (const char *)q2->p + off * 32, h, (char *)d_proofs + (n + i) * PT, pre));
        off += h;
    }
    return SCZ_OK;
}

// d_open: dpoly_comm.rs:355-398.  Leader: value + log2(N) root proofs ++ n summed proofs; others (0, []) (:387).
// *count is known at once (it depends on the role only); the values arrive when `D` has run.
int32_t d_open_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point,
                     size_t npoint, void *d_value, void *d_proofs, size_t *count) {
    bool ok;
    size_t n = log2_exact(len, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "d_open: length %zu is not a power of two", len);
    Net *net = ctx->net;
    const size_t N = net->n_parties, PT = SCZ_G1_JAC_BYTES;
    size_t pl = log2_exact(N, &ok);
    if (!ok) return ctx->fail(SCZ_ERR_NOT_POW2, "d_open: %zu parties is not a power of two", N);
    if (npoint < pl + n) return ctx->fail(SCZ_ERR_BAD_ARG, "d_open: point has %zu coordinates, %zu needed", npoint, pl + n);
    const size_t payload = 32 + n * PT;
    DevTmp *loc = nullptr;
    SCZ_TRY(D.tmp(payload, &loc));
    SCZ_TRY(open_defer(ctx, D, srs, d_peval, len, (const char *)d_point + pl * 32, loc->p, (char *)loc->p + 32));   // :366
    if (count) *count = net->is_leader() ? pl + n : 0;
    Deferred *Dp = &D;
    D.then([=]() -> int32_t {
        DevTmp *recv = nullptr, *lz = nullptr;
        if (net->is_leader()) {
            SCZ_TRY(Dp->tmp(N * payload, &recv));
            SCZ_TRY(Dp->tmp(N * 32, &lz));
        }
        Dp->gather(loc->p, recv ? recv->p : nullptr, payload, 32 + 8 + 48 * n);   // :368
        Dp->then_gathered([=]() -> int32_t {
            // the scatter half of leader_compute_element: the leader keeps (value, proofs), every worker is sent
            // (F::zero(), vec![]) = 32 + 8 serialised bytes (:382-391); only the byte counters see it here
            net->count_scatter(32 + 8);
            if (!net->is_leader()) {
                SCZ_CUDA(ctx, cudaMemsetAsync(d_value, 0, 32, ctx->stream));
                return SCZ_OK;
            }
            k_pick_fr<<<1, 32 * (uint32_t)((N + 31) / 32), 0, ctx->stream>>>(recv->p, payload, (uint32_t)N, lz->p);
            SCZ_LAUNCH_CHECK(ctx);
            SCZ_TRY(open_defer(ctx, *Dp, srs, lz->p, N, d_point, d_value, d_proofs));       // root_open :377 (next flush)
            if (n) Dp->add_colsum(recv->p, payload, 32, (uint32_t)N, (uint32_t)n, (char *)d_proofs + pl * PT, 1);   // :374-376
            return SCZ_OK;
        });
        return SCZ_OK;
    });
    return SCZ_OK;
}

// ---- the queued leader closures of a round, one launch per kind (deferred.h)
__global__ void __launch_bounds__(64) k_g1_colsum_multi(const Deferred::ColsumJob *jobs, uint32_t njobs, uint32_t total_cols) {
    uint32_t t = blockIdx.x * 64 + threadIdx.x;
    if (t >= total_cols) return;
    uint32_t lo = 0, hi = njobs - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (jobs[mid].col_base <= t) lo = mid;
        else hi = mid - 1;
    }
    const Deferred::ColsumJob job = jobs[lo];
    uint32_t i = t - job.col_base;
    G1X acc = G1X::inf();
    for (uint32_t j = 0; j < job.parties; j++) {
        const char *p = reinterpret_cast<const char *>(job.in) + (size_t)j * job.stride + job.off;
        acc = g1x_add(acc, g1x_from_jac(g1j_load(p, i)));
    }
    G1Jac r = g1x_to_jac(acc);
    for (uint32_t k = 0; k < job.replicate; k++) g1j_store(job.out, (size_t)k * job.cols + i, r);
}

// One collective for all gathers registered in this stage.  Without a real net (leader simulator) every request is a
// replicate launch anyway, so they go out one by one.
int32_t Deferred::do_gathers() {
    if (gathers.empty()) return SCZ_OK;
    Net *net = ctx->net;
    std::vector<XferReq> reqs;
    reqs.swap(gathers);
    if (!net->real() || reqs.size() == 1) {
        for (auto &r : reqs) SCZ_TRY(net->gather(ctx, r.send, r.recv, r.bytes, r.wire));
        return SCZ_OK;
    }
    const size_t N = net->n_parties;
    size_t T = 0, wire = 0;
    for (auto &r : reqs) T += r.bytes, wire += r.wire;
    DevTmp sendbuf(ctx), recvbuf(ctx);
    SCZ_TRY(sendbuf.alloc(T));
    if (net->is_leader()) SCZ_TRY(recvbuf.alloc(N * T));
    size_t off = 0;
    for (auto &r : reqs) {
        SCZ_CUDA(ctx, cudaMemcpyAsync((char *)sendbuf.p + off, r.send, r.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        off += r.bytes;
    }
    SCZ_TRY(net->gather(ctx, sendbuf.p, recvbuf.p, T, wire));
    if (net->is_leader()) {
        off = 0;
        for (auto &r : reqs) {   // [party][T] -> the request's own [party][bytes]
            SCZ_CUDA(ctx, cudaMemcpy2DAsync(r.recv, r.bytes, (const char *)recvbuf.p + off, T, r.bytes, N,
                                            cudaMemcpyDeviceToDevice, ctx->stream));
            off += r.bytes;
        }
    }
    return SCZ_OK;
}
int32_t Deferred::do_scatters() {
    if (scatters.empty()) return SCZ_OK;
    Net *net = ctx->net;
    std::vector<XferReq> reqs;
    reqs.swap(scatters);
    if (!net->real() || reqs.size() == 1) {
        for (auto &r : reqs) SCZ_TRY(net->scatter(ctx, r.send, r.recv, r.bytes, r.wire));
        return SCZ_OK;
    }
    const size_t N = net->n_parties;
    size_t T = 0, wire = 0;
    for (auto &r : reqs) T += r.bytes, wire += r.wire;
    DevTmp sendbuf(ctx), recvbuf(ctx);
    SCZ_TRY(recvbuf.alloc(T));
    size_t off = 0;
    if (net->is_leader()) {
        SCZ_TRY(sendbuf.alloc(N * T));
        for (auto &r : reqs) {   // the request's [party][bytes] -> [party][T]
            SCZ_CUDA(ctx, cudaMemcpy2DAsync((char *)sendbuf.p + off, T, r.send, r.bytes, r.bytes, N, cudaMemcpyDeviceToDevice,
                                            ctx->stream));
            off += r.bytes;
        }
    }
    SCZ_TRY(net->scatter(ctx, sendbuf.p, recvbuf.p, T, wire));
    off = 0;
    for (auto &r : reqs) {
        SCZ_CUDA(ctx, cudaMemcpyAsync(r.recv, (const char *)recvbuf.p + off, r.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        off += r.bytes;
    }
    return SCZ_OK;
}

int32_t Deferred::flush_closures() {
    if (!pss_jobs.empty()) {
        ProfScope ps(ctx, SCZ_K_PSS);
        int32_t rc = pss_dmsm_multi(ctx, pss_pp, pss_jobs.data(), pss_jobs.size());
        pss_jobs.clear();
        SCZ_TRY(rc);
    }
    if (!colsum_jobs.empty()) {
        ProfScope ps(ctx, SCZ_K_PSS);
        uint32_t cols = 0;
        for (auto &j : colsum_jobs) {
            j.col_base = cols;
            cols += j.cols;
        }
        DevTmp d(ctx);
        SCZ_TRY(d.alloc(colsum_jobs.size() * sizeof(ColsumJob)));
        SCZ_TRY(ctx->h2d_staged(d.p, colsum_jobs.data(), colsum_jobs.size() * sizeof(ColsumJob)));
        k_g1_colsum_multi<<<ceil_div_u32(cols, 64), 64, 0, ctx->stream>>>(d.as<ColsumJob>(), (uint32_t)colsum_jobs.size(), cols);
        colsum_jobs.clear();
        SCZ_LAUNCH_CHECK(ctx);
    }
    return SCZ_OK;
}

#define SCZ_RUN_NOW(call)        \
    do {                         \
        Deferred D(ctx);         \
        SCZ_TRY(call);           \
        return D.run();          \
    } while (0)
int32_t commit_dev(Ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, void *d_out) {
    SCZ_RUN_NOW(commit_defer(ctx, D, srs, d_peval, len, d_out));
}
int32_t c_commit_dev(Ctx *ctx, const scz_srs *srs, const scz_pp *pp, const void *const *d_pevals, const size_t *lens,
                     size_t batch, void *d_out) {
    SCZ_RUN_NOW(c_commit_defer(ctx, D, srs, pp, d_pevals, lens, batch, d_out));
}
int32_t d_commit_dev(Ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, void *d_out) {
    SCZ_RUN_NOW(d_commit_defer(ctx, D, srs, d_peval, len, d_out));
}
int32_t open_dev(Ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point, void *d_value,
                 void *d_proofs) {
    SCZ_RUN_NOW(open_defer(ctx, D, srs, d_peval, len, d_point, d_value, d_proofs));
}
int32_t c_open_dev(Ctx *ctx, const scz_srs *srs, const scz_pp *pp, const void *d_peval, size_t len, const void *d_point,
                   void *d_value, void *d_proofs) {
    SCZ_RUN_NOW(c_open_defer(ctx, D, srs, pp, d_peval, len, d_point, d_value, d_proofs));
}
int32_t d_open_dev(Ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point, size_t npoint,
                   void *d_value, void *d_proofs, size_t *count) {
    SCZ_RUN_NOW(d_open_defer(ctx, D, srs, d_peval, len, d_point, npoint, d_value, d_proofs, count));
}

}   // namespace scz

using namespace scz;

#define NEED(h, cond, what) \
    if (!(h)) return SCZ_ERR_BAD_ARG; \
    if (!(cond)) return (h)->c.fail(SCZ_ERR_BAD_ARG, what ": null or bad argument")

extern "C" {

int32_t scz_srs_from_device_levels(scz_ctx *h, size_t levels, const void *const *d_levels, const size_t *lens, scz_srs **out) {
    scz::DeviceGuard dg__(h);
    NEED(h, out && (levels == 0 || (d_levels && lens)), "srs");
    scz_srs *s = new scz_srs();
    s->device = h->c.device;
    for (size_t i = 0; i < levels; i++) {
        s->level.push_back(d_levels[i]);
        s->len.push_back(lens[i]);
    }
    *out = s;
    return SCZ_OK;
}
int32_t scz_srs_from_host_levels(scz_ctx *h, size_t levels, const void *const *levels_host, const size_t *lens, scz_srs **out) {
    scz::DeviceGuard dg__(h);
    NEED(h, out && (levels == 0 || (levels_host && lens)), "srs");
    Ctx *c = &h->c;
    scz_srs *s = new scz_srs();
    s->device = c->device;
    for (size_t i = 0; i < levels; i++) {
        void *d = nullptr;
        cudaError_t e = cudaMalloc(&d, lens[i] ? lens[i] * SCZ_G1_AFFINE_BYTES : 1);
        if (e == cudaSuccess && lens[i])
            e = cudaMemcpyAsync(d, levels_host[i], lens[i] * SCZ_G1_AFFINE_BYTES, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) {
            for (void *p : s->owned) cudaFree(p);
            delete s;
            return c->cuda(e, "srs upload");
        }
        s->owned.push_back(d);
        s->level.push_back(d);
        s->len.push_back(lens[i]);
    }
    SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = s;
    return SCZ_OK;
}
void scz_srs_free(scz_srs *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    for (void *p : s->owned) cudaFree(p);
    delete s;
}
int32_t scz_srs_info(const scz_srs *s, size_t *levels) {
    if (!s) return SCZ_ERR_BAD_ARG;
    if (levels) *levels = s->level.size();
    return SCZ_OK;
}

int32_t scz_pss2ss_dev(scz_ctx *h, const scz_pp *pp, const void *d_share, void *d_out) {
    scz::DeviceGuard dg__(h);
    NEED(h, pp && d_share && d_out, "pss2ss");
    return pss2ss_dev(&h->c, pp, d_share, d_out);
}
int32_t scz_degree_reduce_dev(scz_ctx *h, const scz_pp *pp, const void *d_share, void *d_out) {
    scz::DeviceGuard dg__(h);
    NEED(h, pp && d_share && d_out, "degree_reduce");
    return degree_reduce_dev(&h->c, pp, d_share, d_out);
}
int32_t scz_sumcheck_product_dev(scz_ctx *h, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                                 void *d_out) {
    scz::DeviceGuard dg__(h);
    NEED(h, d_f && d_g && d_out && (len <= 1 || d_challenge), "sumcheck_product");
    return sumcheck_product_dev(&h->c, d_f, d_g, len, d_challenge, d_out);
}
int32_t scz_c_sumcheck_product_dev(scz_ctx *h, const scz_pp *pp, const void *d_f, const void *d_g, size_t len,
                                   const void *d_challenge, void *d_out) {
    scz::DeviceGuard dg__(h);
    NEED(h, pp && d_f && d_g && d_out && d_challenge, "c_sumcheck_product");
    return c_sumcheck_product_dev(&h->c, pp, d_f, d_g, len, d_challenge, d_out);
}
int32_t scz_d_sumcheck_product_dev(scz_ctx *h, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                                   void *d_out, size_t *count) {
    scz::DeviceGuard dg__(h);
    NEED(h, d_f && d_g && d_challenge && (d_out || h->c.net->party_id != 0), "d_sumcheck_product");
    return d_sumcheck_product_dev(&h->c, d_f, d_g, len, d_challenge, d_out, count);
}
int32_t scz_sumcheck_dev(scz_ctx *h, const void *d_f, size_t len, const void *d_challenge, void *d_out) {
    scz::DeviceGuard dg__(h);
    NEED(h, d_f && d_out && (len <= 1 || d_challenge), "sumcheck");
    return sumcheck_dev(&h->c, d_f, len, d_challenge, d_out);
}
int32_t scz_c_sumcheck_dev(scz_ctx *h, const scz_pp *pp, const void *d_f, size_t len, const void *d_challenge, void *d_out) {
    scz::DeviceGuard dg__(h);
    NEED(h, pp && d_f && d_out && d_challenge, "c_sumcheck");
    return c_sumcheck_dev(&h->c, pp, d_f, len, d_challenge, d_out);
}
int32_t scz_d_sumcheck_dev(scz_ctx *h, const void *d_f, size_t len, const void *d_challenge, void *d_out, size_t *count) {
    scz::DeviceGuard dg__(h);
    NEED(h, d_f && d_challenge && (d_out || h->c.net->party_id != 0), "d_sumcheck");
    return d_sumcheck_dev(&h->c, d_f, len, d_challenge, d_out, count);
}
int32_t scz_d_acc_product_dev(scz_ctx *h, const void *d_x, size_t m, void *d_subtree, void *d_leader_tree) {
    scz::DeviceGuard dg__(h);
    NEED(h, d_x && d_subtree && (d_leader_tree || h->c.net->party_id != 0), "d_acc_product");
    return d_acc_product_dev(&h->c, d_x, m, d_subtree, d_leader_tree);
}
int32_t scz_commit_dev(scz_ctx *h, const scz_srs *srs, const void *d_peval, size_t len, void *d_out) {
    scz::DeviceGuard dg__(h);
    NEED(h, srs && d_peval && d_out, "commit");
    return commit_dev(&h->c, srs, d_peval, len, d_out);
}
int32_t scz_c_commit_dev(scz_ctx *h, const scz_srs *srs, const scz_pp *pp, const void *const *d_pevals, const size_t *lens,
                         size_t batch, void *d_out) {
    scz::DeviceGuard dg__(h);
    NEED(h, srs && pp && (batch == 0 || (d_pevals && lens && d_out)), "c_commit");
    return c_commit_dev(&h->c, srs, pp, d_pevals, lens, batch, d_out);
}
int32_t scz_d_commit_dev(scz_ctx *h, const scz_srs *srs, const void *d_peval, size_t len, void *d_out) {
    scz::DeviceGuard dg__(h);
    NEED(h, srs && d_peval && d_out, "d_commit");
    return d_commit_dev(&h->c, srs, d_peval, len, d_out);
}
int32_t scz_open_dev(scz_ctx *h, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point, void *d_value,
                     void *d_proofs) {
    scz::DeviceGuard dg__(h);
    NEED(h, srs && d_peval && d_value && (len <= 1 || (d_point && d_proofs)), "open");
    return open_dev(&h->c, srs, d_peval, len, d_point, d_value, d_proofs);
}
int32_t scz_c_open_dev(scz_ctx *h, const scz_srs *srs, const scz_pp *pp, const void *d_peval, size_t len,
                       const void *d_point, void *d_value, void *d_proofs) {
    scz::DeviceGuard dg__(h);
    NEED(h, srs && pp && d_peval && d_value && d_point && d_proofs, "c_open");
    return c_open_dev(&h->c, srs, pp, d_peval, len, d_point, d_value, d_proofs);
}
int32_t scz_d_open_dev(scz_ctx *h, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point, size_t npoint,
                       void *d_value, void *d_proofs, size_t *count) {
    scz::DeviceGuard dg__(h);
    NEED(h, srs && d_peval && d_value && d_point && (d_proofs || h->c.net->party_id != 0), "d_open");
    return d_open_dev(&h->c, srs, d_peval, len, d_point, npoint, d_value, d_proofs, count);
}

}   // extern "C"
