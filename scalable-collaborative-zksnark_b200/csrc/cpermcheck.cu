// The collaborative (PSS) permutation check and its masked product accumulation:
//   d_unpack2_many             dist-primitive/src/unpack.rs:55-70
//   degree_reduce_many         dist-primitive/src/degree_reduce.rs:10-26
//   c_acc_product              dist-primitive/src/dacc_product.rs:296-363
//   merge                      dist-primitive/src/dacc_product.rs:416-428
//   c_acc_product_and_share    dist-primitive/src/dacc_product.rs:66-292
//   cpermcheck                 hyperplonk/src/dhyperplonk.rs:1249-1385
// This is the paper's baseline prodcheck (N hub rounds with a moving hub).  The reference marks parts of it as
// "NOTE: We do not guarantee correctness here" (dacc_product.rs:353): it is a cost skeleton, and it is restated
// here statement by statement -- including what its build without `comm` substitutes for received data
// (:186-202) -- not repaired.
#include <vector>

#include "field.cuh"
#include "net.h"
#include "protocols.h"

namespace scz {

static inline size_t ilog2c(size_t v) {
    size_t l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}

__global__ void __launch_bounds__(256) k_fr_mul_inplace(void *a, const void *b, size_t n) {
    size_t i = blockIdx.x * (size_t)256 + threadIdx.x;
    if (i >= n) return;
    fp_store<FrP>(a, i, fp_mul(fp_load_rw<FrP>(a, i), fp_load<FrP>(b, i)));
}
__global__ void __launch_bounds__(256) k_fr_mul(const void *a, const void *b, void *out, size_t n) {
    size_t i = blockIdx.x * (size_t)256 + threadIdx.x;
    if (i >= n) return;
    fp_store<FrP>(out, i, fp_mul(fp_load_rw<FrP>(a, i), fp_load_rw<FrP>(b, i)));
}
// leader tree top (dacc_product.rs:354-358): lt[i] = lt[x0] * lt[x1] for i in [first, last), (x0, x1) = sub_index(i);
// lt[last] = 0.  The products depend on each other: one thread.
__global__ void k_leader_tree_top(void *lt, uint32_t first, uint32_t last) {
    if (threadIdx.x || blockIdx.x) return;
    for (uint32_t i = first; i < last; i++) {
        uint32_t top = 31 - __clz(i);
        uint32_t x = (i & ~(1u << top)) << 1;
        fp_store<FrP>(lt, i, fp_mul(fp_load_rw<FrP>(lt, x), fp_load_rw<FrP>(lt, x + 1)));
    }
    fp_store<FrP>(lt, last, Fr::zero());
}

// d_unpack2_many (unpack.rs:55-70): every party sends `len` shares to `receiver`; the receiver transposes and
// unpack2's every position: d_out gets len * l values, position-major.  Others get nothing (*got = false).
static int32_t d_unpack2_many(Ctx *ctx, const scz_pp *pp, const void *d_share, size_t len, uint32_t receiver, void *d_out,
                              bool *got) {
    Net *net = ctx->net;
    const size_t N = net->n_parties;
    const bool me = receiver == net->party_id;
    DevTmp recv(ctx);
    if (me) SCZ_TRY(recv.alloc(N * len * 32));
    SCZ_TRY(net->gather_to(ctx, receiver, d_share, recv.p, len * 32, 8 + 32 * len));
    *got = me;
    if (!me) return SCZ_OK;
    ProfScope ps(ctx, SCZ_K_PSS);
    return pss_apply(ctx, pp, PSS_UNPACK2, 0, recv.p, N, 1, len, len, d_out, pp->l, 1);
}

// degree_reduce_many (degree_reduce.rs:10-26): gather; leader: per position unpack2 + pack_from_public; scatter
static int32_t degree_reduce_many(Ctx *ctx, const scz_pp *pp, const void *d_shares, size_t len, void *d_out) {
    Net *net = ctx->net;
    const size_t N = net->n_parties, l = pp->l;
    DevTmp recv(ctx), sec(ctx), send(ctx);
    if (net->is_leader()) {
        SCZ_TRY(recv.alloc(N * len * 32));
        SCZ_TRY(sec.alloc(len * l * 32));
        SCZ_TRY(send.alloc(N * len * 32));
    }
    SCZ_TRY(net->gather(ctx, d_shares, recv.p, len * 32, 8 + 32 * len));
    if (net->is_leader()) {
        ProfScope ps(ctx, SCZ_K_PSS);
        SCZ_TRY(pss_apply(ctx, pp, PSS_UNPACK2, 0, recv.p, N, 1, len, len, sec.p, l, 1));
        SCZ_TRY(pss_apply(ctx, pp, PSS_PACK, 0, sec.p, l, l, 1, len, send.p, 1, len));
    }
    return net->scatter(ctx, send.p, d_out, len * 32, 8 + 32 * len);
}

// merge (dacc_product.rs:416-428): results is [N][r]; level-order interleave: take 2^k entries of every row, then
// 2^(k-1), ... while they fit.  Returns the number of entries written.
static size_t merge_rows(Ctx *ctx, const void *d_results, size_t N, size_t r, void *d_out, int32_t *rc) {
    size_t num = 1;
    while (num < r + 1) num <<= 1;   // (r + 1).next_power_of_two()
    num >>= 1;
    size_t start = 0, written = 0;
    *rc = SCZ_OK;
    while (num && start + num <= r) {
        cudaError_t e = cudaMemcpy2DAsync((char *)d_out + written * 32, num * 32, (const char *)d_results + start * 32, r * 32,
                                          num * 32, N, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) {
            *rc = ctx->cuda(e, "merge");
            return written;
        }
        written += num * N;
        start += num;
        num >>= 1;
    }
    return written;
}

// c_acc_product (dacc_product.rs:296-363): subtree (2m entries) everywhere; leader tree (N * N entries) on the leader
static int32_t c_acc_product(Ctx *ctx, const scz_pp *pp, const void *d_x, size_t m, void *d_subtree, void *d_leader_tree) {
    Net *net = ctx->net;
    const size_t N = pp->n;
    if (2 * m < N) return ctx->fail(SCZ_ERR_BAD_ARG, "c_acc_product: subtree of %zu entries is shorter than the %zu parties", 2 * m, N);
    SCZ_TRY(acc_product_tree(ctx, d_x, m, d_subtree));                                       // :305-313
    DevTmp recv(ctx);
    if (net->is_leader()) SCZ_TRY(recv.alloc(N * N * 32));
    SCZ_TRY(net->gather(ctx, (const char *)d_subtree + (2 * m - N) * 32, recv.p, N * 32, 8 + 32 * N));   // last N entries :320-328
    if (!net->is_leader()) return SCZ_OK;
    // level-order merge of the N received tails (:338-349): N/2, N/4, ..., 1 entries of every party
    size_t written = 0, start = 0;
    for (size_t layer = N >> 1; layer > 0; layer >>= 1) {
        SCZ_CUDA(ctx, cudaMemcpy2DAsync((char *)d_leader_tree + written * 32, layer * 32, (const char *)recv.p + start * 32, N * 32,
                                        layer * 32, N, cudaMemcpyDeviceToDevice, ctx->stream));
        written += layer * N;
        start += layer;
    }
    k_leader_tree_top<<<1, 32, 0, ctx->stream>>>(d_leader_tree, (uint32_t)(N * N - N), (uint32_t)(N * N - 1));   // :354-358
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}

// c_acc_product_and_share (dacc_product.rs:66-292).  shares / masks / unmask*: L entries each; outputs: L entries each.
int32_t c_acc_product_and_share_dev(Ctx *ctx, const scz_pp *pp, const void *d_shares, const void *d_masks, const void *d_unmask0,
                                    const void *d_unmask1, const void *d_unmask2, size_t L, void *d_share0, void *d_share1,
                                    void *d_share2) {
    Net *net = ctx->net;
    const size_t N = pp->n, l = pp->l;
    if (net->n_parties != N) return ctx->fail(SCZ_ERR_BAD_ARG, "c_acc_product_and_share: %u parties but pp.n = %zu", net->n_parties, N);
    if (!(L > N) || L % N) return ctx->fail(SCZ_ERR_BAD_ARG, "c_acc_product_and_share: %zu shares for %zu parties", L, N);   // :82
    const size_t block = L / N, m = block * l;
    if ((m & (m - 1)) || m < N) return ctx->fail(SCZ_ERR_NOT_POW2, "c_acc_product_and_share: %zu values per party", m);
    cudaStream_t st = ctx->stream;
    const uint32_t me = net->party_id;

    // masked x (:88-108): every party ends with the plain masked values of ITS block
    DevTmp masked(ctx), masked_x(ctx);
    SCZ_TRY(masked.alloc(L * 32));
    SCZ_TRY(masked_x.alloc(m * 32));
    {
        ProfScope ps(ctx, SCZ_K_POINTWISE);
        k_fr_mul<<<ceil_div_u32(L, 256), 256, 0, st>>>(d_shares, d_masks, masked.p, L);
        SCZ_LAUNCH_CHECK(ctx);
    }
    for (size_t i = 0; i < N; i++) {
        bool got = false;
        SCZ_TRY(d_unpack2_many(ctx, pp, (const char *)masked.p + i * block * 32, block, (uint32_t)i, masked_x.p, &got));
    }

    // tree (:111-113)
    DevTmp subtree(ctx), ltree(ctx);
    SCZ_TRY(subtree.alloc(2 * m * 32));
    SCZ_TRY(ltree.alloc(N * N * 32));
    SCZ_TRY(c_acc_product(ctx, pp, masked_x.p, m, subtree.p, ltree.p));

    // share matrices of the subtree (:115-152): [party][chunk]
    const size_t r0 = (m - N / 2) / l, r2 = (m - N) / l;
    DevTmp sm0(ctx), sm1(ctx), sm2(ctx), res0(ctx), res1(ctx), res2(ctx);
    SCZ_TRY(sm0.alloc((r0 ? N * r0 : 1) * 32));
    SCZ_TRY(sm1.alloc((r0 ? N * r0 : 1) * 32));
    SCZ_TRY(sm2.alloc((r2 ? N * r2 : 1) * 32));
    {
        ProfScope ps(ctx, SCZ_K_PSS);
        if (r0) {
            SCZ_TRY(pss_apply(ctx, pp, PSS_PACK, 0, subtree.p, l, 2 * l, 2, r0, sm0.p, 1, r0));                     // even entries
            SCZ_TRY(pss_apply(ctx, pp, PSS_PACK, 0, (const char *)subtree.p + 32, l, 2 * l, 2, r0, sm1.p, 1, r0));   // odd entries
        }
        if (r2) SCZ_TRY(pss_apply(ctx, pp, PSS_PACK, 0, (const char *)subtree.p + m * 32, l, l, 1, r2, sm2.p, 1, r2));   // skip(len/2)
    }
    // N hub rounds (:154-203): hub i sends row j of its three matrices to party j
    const void *rs0 = sm0.p, *rs1 = sm1.p, *rs2 = sm2.p;
    if (net->real()) {
        SCZ_TRY(res0.alloc((r0 ? N * r0 : 1) * 32));
        SCZ_TRY(res1.alloc((r0 ? N * r0 : 1) * 32));
        SCZ_TRY(res2.alloc((r2 ? N * r2 : 1) * 32));
        for (size_t i = 0; i < N; i++) {
            bool got;
            SCZ_TRY(net->scatter_from(ctx, (uint32_t)i, sm0.p, (char *)res0.p + i * r0 * 32, r0 * 32, 8 + 32 * r0, &got));
            SCZ_TRY(net->scatter_from(ctx, (uint32_t)i, sm1.p, (char *)res1.p + i * r0 * 32, r0 * 32, 8 + 32 * r0, &got));
            SCZ_TRY(net->scatter_from(ctx, (uint32_t)i, sm2.p, (char *)res2.p + i * r2 * 32, r2 * 32, 8 + 32 * r2, &got));
        }
        rs0 = res0.p, rs1 = res1.p, rs2 = res2.p;
    } else {
        // build without `comm` (:196-202): the received value is dropped and row i of the OWN matrices stands in
        DevTmp sink(ctx);
        SCZ_TRY(sink.alloc((r0 > r2 ? r0 : r2) * 32 + 32));
        for (size_t i = 0; i < N; i++) {
            bool got;
            SCZ_TRY(net->scatter_from(ctx, (uint32_t)i, sm0.p, sink.p, r0 * 32, 8 + 32 * r0, &got));
            SCZ_TRY(net->scatter_from(ctx, (uint32_t)i, sm1.p, sink.p, r0 * 32, 8 + 32 * r0, &got));
            SCZ_TRY(net->scatter_from(ctx, (uint32_t)i, sm2.p, sink.p, r2 * 32, 8 + 32 * r2, &got));
        }
    }
    int32_t rc;
    size_t n0 = merge_rows(ctx, rs0, N, r0, d_share0, &rc);                                   // :204-209
    SCZ_TRY(rc);
    size_t n1 = merge_rows(ctx, rs1, N, r0, d_share1, &rc);
    SCZ_TRY(rc);
    size_t n2 = merge_rows(ctx, rs2, N, r2, d_share2, &rc);
    SCZ_TRY(rc);

    // the leader shares its tree (:212-262): even / odd entries and -- for v(1,x) -- the WHOLE tree (:244-249)
    const size_t q0 = N * N / 2 / l, q2 = N * N / l;
    if (n0 + q0 != L || n1 + q0 != L || n2 + q2 != L)
        return ctx->fail(SCZ_ERR_BAD_ARG, "c_acc_product_and_share: merged lengths %zu/%zu/%zu + %zu/%zu do not give %zu", n0, n1, n2, q0, q2, L);
    DevTmp lm0(ctx), lm1(ctx), lm2(ctx);
    if (net->is_leader()) {
        SCZ_TRY(lm0.alloc(N * q0 * 32));
        SCZ_TRY(lm1.alloc(N * q0 * 32));
        SCZ_TRY(lm2.alloc(N * q2 * 32));
        ProfScope ps(ctx, SCZ_K_PSS);
        SCZ_TRY(pss_apply(ctx, pp, PSS_PACK, 0, ltree.p, l, 2 * l, 2, q0, lm0.p, 1, q0));
        SCZ_TRY(pss_apply(ctx, pp, PSS_PACK, 0, (const char *)ltree.p + 32, l, 2 * l, 2, q0, lm1.p, 1, q0));
        SCZ_TRY(pss_apply(ctx, pp, PSS_PACK, 0, ltree.p, l, l, 1, q2, lm2.p, 1, q2));
    }
    SCZ_TRY(net->scatter(ctx, lm0.p, (char *)d_share0 + n0 * 32, q0 * 32, 8 + 32 * q0));
    SCZ_TRY(net->scatter(ctx, lm1.p, (char *)d_share1 + n1 * 32, q0 * 32, 8 + 32 * q0));
    SCZ_TRY(net->scatter(ctx, lm2.p, (char *)d_share2 + n2 * 32, q2 * 32, 8 + 32 * q2));

    // unmask (:265-275)
    {
        ProfScope ps(ctx, SCZ_K_POINTWISE);
        uint32_t g = ceil_div_u32(L, 256);
        k_fr_mul_inplace<<<g, 256, 0, st>>>(d_share0, d_unmask0, L);
        SCZ_LAUNCH_CHECK(ctx);
        k_fr_mul_inplace<<<g, 256, 0, st>>>(d_share1, d_unmask1, L);
        SCZ_LAUNCH_CHECK(ctx);
        k_fr_mul_inplace<<<g, 256, 0, st>>>(d_share2, d_unmask2, L);
        SCZ_LAUNCH_CHECK(ctx);
    }
    // "These three shares need to be reduced ... we run 1/N of it" (:278-285): the results are dropped by the reference
    {
        size_t red = L / N * 2;
        DevTmp drop(ctx);
        SCZ_TRY(drop.alloc(red * 32));
        SCZ_TRY(degree_reduce_many(ctx, pp, d_share0, red, drop.p));
        SCZ_TRY(degree_reduce_many(ctx, pp, d_share1, red, drop.p));
        SCZ_TRY(degree_reduce_many(ctx, pp, d_share2, red, drop.p));
    }
    (void)me;
    return SCZ_OK;
}

// cpermcheck (dhyperplonk.rs:1249-1385).  Tables of L = 4 * 2^n / l entries.
struct HpOutC {
    Ctx *ctx;
    char *tri, *pts, *val;
    size_t tri_cap, pts_cap, val_cap, items_cap;
    size_t tri_n = 0, pts_n = 0, val_n = 0, items_n = 0;
    scz_hp_item *items;
    int32_t reserve(size_t t, size_t p) {
        if (items_n >= items_cap || tri_n + t > tri_cap || pts_n + p > pts_cap || val_n + 1 > val_cap)
            return ctx->fail(SCZ_ERR_BAD_ARG, "cpermcheck: output arenas too small (see scz_dhyperplonk_sizes)");
        return SCZ_OK;
    }
    void push(uint32_t kind, size_t t, size_t p, size_t v) {
        scz_hp_item &it = items[items_n++];
        it.kind = kind;
        it.triples_off = (uint32_t)tri_n, it.triples_cnt = (uint32_t)t;
        it.points_off = (uint32_t)pts_n, it.points_cnt = (uint32_t)p;
        it.value_off = (uint32_t)val_n, it.value_cnt = (uint32_t)v;
        tri_n += t, pts_n += p, val_n += v;
    }
};

int32_t cpermcheck_dev(Ctx *ctx, size_t n, const scz_cperm_pk *pk, const scz_pp *pp, HpOutC &o) {
    const size_t l = pp->l, ll = ilog2c(l);
    const size_t L = ((size_t)4 << n) / l;     // gate_count * 4 with gate_count = 2^n / l (:1270)
    const size_t PT = SCZ_G1_JAC_BYTES;
    Deferred D(ctx);
    DevTmp num(ctx), den(ctx);
    SCZ_TRY(num.alloc(L * 32));
    SCZ_TRY(den.alloc(L * 32));
    SCZ_TRY(fr_pointwise(ctx, 2, pk->V, pk->sid, pk->alpha_beta, num.p, L));          // :1278-1280
    SCZ_TRY(fr_pointwise(ctx, 2, pk->eq_r1, pk->ssigma, pk->alpha_beta, den.p, L));   // :1281-1283
    const size_t nopen = ilog2c(L) + ll, ntri = ilog2c(L) + ll + 1;
    auto commit = [&](const void *tab) -> int32_t {
        SCZ_TRY(o.reserve(0, 1));
        SCZ_TRY(c_commit_defer(ctx, D, pk->c_commitment, pp, &tab, &L, 1, o.pts + o.pts_n * PT));
        o.push(SCZ_HP_WIRING_COMMIT, 0, 1, 0);
        return SCZ_OK;
    };
    auto open = [&](const void *tab) -> int32_t {
        SCZ_TRY(o.reserve(0, nopen));
        SCZ_TRY(c_open_defer(ctx, D, pk->c_commitment, pp, tab, L, pk->challenge_r1, o.val + o.val_n * 32, o.pts + o.pts_n * PT));
        o.push(SCZ_HP_WIRING_OPEN, 0, nopen, 1);
        return SCZ_OK;
    };
    auto sumcheck = [&](const void *f, const void *g) -> int32_t {
        SCZ_TRY(o.reserve(ntri, 0));
        SCZ_TRY(c_sumcheck_product_dev(ctx, pp, f, g, L, pk->challenge_r1, o.tri + o.tri_n * SCZ_TRIPLE_BYTES));
        o.push(SCZ_HP_WIRING_PROOF, ntri, 0, 0);
        return SCZ_OK;
    };
    SCZ_TRY(commit(pk->ssigma));                                                       // :1289-1310
    SCZ_TRY(open(pk->ssigma));
    SCZ_TRY(commit(pk->sid));
    SCZ_TRY(open(pk->sid));
    const void *fs[2] = {num.p, den.p};
    for (int k = 0; k < 2; k++) {                                                      // :1311-1376
        DevTmp *v = nullptr;   // vx0 | vx1 | v1x: read by queued MSMs, so owned by D
        SCZ_TRY(D.tmp(3 * L * 32, &v));
        void *vx0 = v->p, *vx1 = (char *)v->p + L * 32, *v1x = (char *)v->p + 2 * L * 32;
        SCZ_TRY(c_acc_product_and_share_dev(ctx, pp, fs[k], pk->mask, pk->unmask0, pk->unmask1, pk->unmask2, L, vx0, vx1, v1x));
        SCZ_TRY(commit(fs[k]));
        SCZ_TRY(open(fs[k]));
        SCZ_TRY(commit(vx0));
        SCZ_TRY(open(vx0));
        SCZ_TRY(commit(vx1));
        SCZ_TRY(open(vx1));
        SCZ_TRY(commit(v1x));
        SCZ_TRY(open(v1x));
        SCZ_TRY(sumcheck(pk->eq_r1, v1x));                                             // :1364-1368
        SCZ_TRY(sumcheck(pk->eq_r1, vx0));
        SCZ_TRY(sumcheck(vx0, vx1));
        SCZ_TRY(open(fs[k]));                                                          // :1370-1374
    }
    return D.run();
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_c_acc_product_and_share_dev(scz_ctx *h, const scz_pp *pp, const void *d_shares, const void *d_masks,
                                        const void *d_unmask0, const void *d_unmask1, const void *d_unmask2, size_t len,
                                        void *d_share0, void *d_share1, void *d_share2) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (!pp || !d_shares || !d_masks || !d_unmask0 || !d_unmask1 || !d_unmask2 || !d_share0 || !d_share1 || !d_share2)
        return h->c.fail(SCZ_ERR_BAD_ARG, "c_acc_product_and_share: null argument");
    return c_acc_product_and_share_dev(&h->c, pp, d_shares, d_masks, d_unmask0, d_unmask1, d_unmask2, len, d_share0, d_share1,
                                       d_share2);
}

int32_t scz_cpermcheck_dev(scz_ctx *h, size_t n, const scz_cperm_pk *pk, const scz_pp *pp, void *d_triples, size_t triples_cap,
                           void *d_points, size_t points_cap, void *d_values, size_t values_cap, scz_hp_item *items,
                           size_t items_cap, size_t *n_items) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (!pk || !pp || !d_triples || !d_points || !d_values || !items || !n_items)
        return c->fail(SCZ_ERR_BAD_ARG, "cpermcheck: null argument");
    const void *need[] = {pk->V, pk->sid, pk->ssigma, pk->eq_r1, pk->mask, pk->unmask0, pk->unmask1, pk->unmask2,
                          pk->challenge_r1, pk->alpha_beta, pk->c_commitment};
    for (const void *p : need)
        if (!p) return c->fail(SCZ_ERR_BAD_ARG, "cpermcheck: a field of scz_cperm_pk is null");
    if (n < 1 || n > 28) return c->fail(SCZ_ERR_BAD_ARG, "cpermcheck: n = %zu", n);
    HpOutC o;
    o.ctx = c;
    o.tri = (char *)d_triples, o.pts = (char *)d_points, o.val = (char *)d_values;
    o.tri_cap = triples_cap, o.pts_cap = points_cap, o.val_cap = values_cap, o.items_cap = items_cap;
    o.items = items;
    int32_t rc = cpermcheck_dev(c, n, pk, pp, o);
    *n_items = o.items_n;
    return rc;
}

}   // extern "C"
