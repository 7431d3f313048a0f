// The collaborative HyperPlonk prover: hyperplonk/src/dhyperplonk.rs:159-571, from after net.sync() (:193)
// to the return -- the region the reference's "Distributed HyperPlonk" timer covers (:194, :561).
//
// This file is only the SCHEDULE: every step is one of the protocol functions of protocols.cu / dmsm.cu /
// poly.cu working on this party's device-resident tables.  Nothing returns to the host between steps except
// through the host-side collectives of a real net; in leader mode the whole proof is one stream of launches.
// Outputs are written straight into three device arenas (triples, points, values); `items` tells the host which
// slice is which entry of the reference's return tuple (:567-570), in the reference's push order.
#include <cstdlib>
#include <functional>
#include <vector>

#include "net.h"
#include "protocols.h"

namespace scz {

static inline size_t ilog2(size_t v) {
    size_t l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}

struct HpOut {
    Ctx *ctx;
    char *tri, *pts, *val;
    size_t tri_cap, pts_cap, val_cap, items_cap;
    size_t tri_n = 0, pts_n = 0, val_n = 0, items_n = 0;
    scz_hp_item *items;

    void *tri_at() const { return tri + tri_n * SCZ_TRIPLE_BYTES; }
    void *pts_at(size_t extra = 0) const { return pts + (pts_n + extra) * SCZ_G1_JAC_BYTES; }
    void *val_at() const { return val + val_n * SCZ_FR_BYTES; }
    // room for one more entry of at most (t triples, p points, one value)?
    int32_t reserve(size_t t, size_t p) {
        if (items_n >= items_cap || tri_n + t > tri_cap || pts_n + p > pts_cap || val_n + 1 > val_cap)
            return ctx->fail(SCZ_ERR_BAD_ARG, "dhyperplonk: output arenas too small (see scz_dhyperplonk_sizes)");
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

// variant: the three provers of hyperplonk/src/dhyperplonk.rs share one schedule
//   HP_FULL           dhyperplonk                (:159-571)
//   HP_DATA_PARALLEL  dhyperplonk_data_parallel  (:573-960): `s` is an input (pk->local_s holds 4gc/l entries, :603)
//                     instead of the all-gather of step 2.a -- nothing else differs
//   HP_PERMCHECK      dpermcheck                 (:962-1247): step 2 alone; returns the wiring triple only
enum HpVariant { HP_FULL = 0, HP_DATA_PARALLEL = 1, HP_PERMCHECK = 2 };

int32_t dhyperplonk_dev(Ctx *ctx, size_t n, const scz_hp_pk *pk, const scz_pp *pp, HpOut &o, HpVariant variant) {
    Net *net = ctx->net;
    const size_t N = net->n_parties, l = pp->l, ll = ilog2(l), s = ilog2(N);
    if (N != pp->n) return ctx->fail(SCZ_ERR_BAD_ARG, "dhyperplonk: %zu parties but pp.n = %zu", N, pp->n);
    if (((size_t)1 << s) != N || n < s + 1 || n > 28)
        return ctx->fail(SCZ_ERR_BAD_ARG, "dhyperplonk: needs 2^k parties and log2(N) < n <= 28 (n = %zu, N = %zu)", n, N);
    const size_t gc = (size_t)1 << n;
    const size_t share_len = gc / l;          // a, b, c, I, S1, S2, eq             (:70-80, :93)
    const size_t v_len = gc * 4 / l;          // V, s                               (:70, :270)
    const size_t slice_len = gc / N;          // I_p, S1_p, S2_p                    (:77-82)
    const size_t hl = gc * 4 / N;             // local_s_p, sid_p, ssigma_p, eq_r*_p, h_p  (:85-98, :324)
    const size_t ls_len = gc * 4 / N / l;     // local_s                            (:189)
    const size_t PT = SCZ_G1_JAC_BYTES;
    const bool leader = net->is_leader();
    const size_t nc = ilog2(share_len) + ll + 1;   // triples of a c_sumcheck_product on 2^n/l shares
    cudaStream_t st = ctx->stream;
    Deferred D(ctx);   // every MSM of the proof is queued here and runs in (at most a few) batched launch sequences
    // early starts of the queued MSM work under the rest of the protocol phase (deferred.h, SCZ_MSM_STREAM=1 only);
    // SCZ_MSM_EARLY is a dev knob: bit 0 after step 1, bit 1 after 2.d, bit 2 after the commits / openings of 2.e.
    // Measured at 2^20 (profiles/r1_pipelined.txt): one early start at the last point is best (167.6 ms per proof vs
    // 173.2 without); more starts cost more in per-sequence overhead than they hide (171.4 / 174.0 ms with two / three)
    static const unsigned early_mask = [] { const char *e = getenv("SCZ_MSM_EARLY"); return e ? (unsigned)strtoul(e, nullptr, 0) : 4u; }();
    auto early = [&](unsigned bit) -> int32_t { return (early_mask >> bit) & 1 ? D.flush_early() : SCZ_OK; };

    // ---- Step 1: commit (:196-217).  The six commitments leave with the openings at the very end (:518-553).
    DevTmp coms(ctx);
    SCZ_TRY(coms.alloc(6 * PT));
    if (variant != HP_PERMCHECK) {
        const void *tabs[3] = {pk->a_evals, pk->b_evals, pk->c_evals};
        for (int k = 0; k < 3; k++)   // three separate c_commit calls, one leader round each (:198-212)
            SCZ_TRY(c_commit_defer(ctx, D, pk->c_commitment, pp, &tabs[k], &share_len, 1, (char *)coms.p + k * PT));
        const void *slc[3] = {pk->I_p, pk->S1_p, pk->S2_p};
        for (int k = 0; k < 3; k++)   // :213-215
            SCZ_TRY(d_commit_defer(ctx, D, pk->d_commitment, slc[k], slice_len, (char *)coms.p + (3 + k) * PT));
        SCZ_TRY(early(0));
    }

    // ---- Step 3: gate identity (:222-260): six collaborative product sumchecks on 2^n/l shares.
    // Sumchecks feed no MSM: their output slots are claimed here, in the reference's push order, but the kernels (and
    // the pss2ss rounds inside) are issued after the early MSM start below, where they hide under the bucket kernel
    // instead of delaying it (`later`; every party defers the same calls, so the collectives still pair up).
    std::vector<std::function<int32_t()>> later;
    DevTmp tmp(ctx), sv(ctx);
    if (variant != HP_PERMCHECK) {
        SCZ_TRY(tmp.alloc(share_len * 32));
        auto c_sum = [&](const void *f, const void *g) -> int32_t {
            SCZ_TRY(o.reserve(nc, 0));
            void *dst = o.tri_at();
            o.push(SCZ_HP_GATE_PROOF, nc, 0, 0);
            later.push_back([=]() -> int32_t { return c_sumcheck_product_dev(ctx, pp, f, g, share_len, pk->challenge, dst); });
            return SCZ_OK;
        };
        auto pointwise = [&](int mode, const void *a, const void *b) {
            void *dst = tmp.p;
            later.push_back([=]() -> int32_t { return fr_pointwise(ctx, mode, a, b, nullptr, dst, share_len); });
        };
        SCZ_TRY(c_sum(pk->eq, pk->S1));                                                     // :230-231
        pointwise(0, pk->a_evals, pk->b_evals);                                             // sum_ab :233-238
        SCZ_TRY(c_sum(pk->S1, tmp.p));                                                      // :240-241
        SCZ_TRY(c_sum(pk->eq, pk->S2));                                                     // :243-244
        SCZ_TRY(c_sum(pk->a_evals, pk->b_evals));                                           // :245-246
        SCZ_TRY(c_sum(pk->S2, pk->a_evals));                                                // :247-248
        pointwise(1, pk->c_evals, pk->I);                                                   // -c + I :251-256
        SCZ_TRY(c_sum(pk->eq, tmp.p));                                                      // :258-259
    }

    // ---- Step 2: wiring identity (:263-513)
    auto d_commit = [&](const void *tab, size_t len) -> int32_t {
        SCZ_TRY(o.reserve(0, 1));
        SCZ_TRY(d_commit_defer(ctx, D, pk->d_commitment, tab, len, o.pts_at()));
        o.push(SCZ_HP_WIRING_COMMIT, 0, 1, 0);
        return SCZ_OK;
    };
    auto d_open = [&](uint32_t kind, const void *tab, size_t len, const void *point, size_t npoint, size_t lead) -> int32_t {
        // `lead` points precede the proofs (the commitment of a gate_identity_commitments entry)
        size_t cap = s + ilog2(len), cnt = 0;
        SCZ_TRY(o.reserve(0, lead + cap));
        SCZ_TRY(d_open_defer(ctx, D, pk->d_commitment, tab, len, point, npoint, o.val_at(), o.pts_at(lead), &cnt));
        o.push(kind, 0, lead + cnt, 1);
        return SCZ_OK;
    };
    auto c_open = [&](uint32_t kind, const void *tab, size_t len, const void *point, size_t lead) -> int32_t {
        size_t cnt = ilog2(len) + ll;
        SCZ_TRY(o.reserve(0, lead + cnt));
        SCZ_TRY(c_open_defer(ctx, D, pk->c_commitment, pp, tab, len, point, o.val_at(), o.pts_at(lead)));
        o.push(kind, 0, lead + cnt, 1);
        return SCZ_OK;
    };
    auto d_sum = [&](const void *f, const void *g, size_t len, const void *challenge) -> int32_t {
        size_t cap = ilog2(len) + s, cnt = 0;
        SCZ_TRY(o.reserve(cap, 0));
        SCZ_TRY(d_sumcheck_product_dev(ctx, f, g, len, challenge, o.tri_at(), &cnt));
        o.push(SCZ_HP_WIRING_PROOF, cnt, 0, 0);
        return SCZ_OK;
    };
    const char *r2 = (const char *)pk->challenge_r2;
    {
        // 2.a (:270-294): N hub rounds in which hub i sends its local_s to everybody = one all-gather;
        // s = local_s^(0) | ... | local_s^(N-1).  The leader simulator repeats the own vector N times (:289-293).
        SCZ_TRY(sv.alloc(v_len * 32));
        if (variant == HP_DATA_PARALLEL)   // s drawn locally (:603): no exchange
            SCZ_CUDA(ctx, cudaMemcpyAsync(sv.p, pk->local_s, v_len * 32, cudaMemcpyDeviceToDevice, st));
        else
            SCZ_TRY(net->all_gather(ctx, pk->local_s, sv.p, ls_len * 32, 8 + 32 * ls_len));
        SCZ_TRY(d_commit(pk->local_s_p, hl));                                               // 2.b :297-302
        {                                                                                    // 2.c :304
            size_t cnt = ilog2(v_len) + ll + 1;
            SCZ_TRY(o.reserve(cnt, 0));
            void *dst = o.tri_at();
            const void *svp = sv.p;
            o.push(SCZ_HP_WIRING_PROOF, cnt, 0, 0);
            later.push_back([=]() -> int32_t { return c_sumcheck_product_dev(ctx, pp, svp, pk->V, v_len, pk->challenge_r1, dst); });
        }
    }
    SCZ_TRY(c_open(SCZ_HP_WIRING_OPEN, pk->V, v_len, pk->challenge_r1, 0));                  // 2.d :306-320
    SCZ_TRY(c_open(SCZ_HP_WIRING_OPEN, pk->V, v_len, pk->challenge_r2, 0));
    SCZ_TRY(d_open(SCZ_HP_WIRING_OPEN, pk->local_s_p, hl, r2, n + 2, 0));
    SCZ_TRY(early(1));   // the two openings of V are a third of the proof's MSM work

    // 2.e (:324-342): num, den, h = num / den, product tree of h
    DevTmp num(ctx), den(ctx), h_p(ctx), subtree(ctx), vx0(ctx), vx1(ctx), ltree(ctx);
    SCZ_TRY(num.alloc(hl * 32));
    SCZ_TRY(den.alloc(hl * 32));
    SCZ_TRY(h_p.alloc(hl * 32));
    SCZ_TRY(subtree.alloc(2 * hl * 32));
    SCZ_TRY(vx0.alloc(hl * 32));
    SCZ_TRY(vx1.alloc(hl * 32));
    SCZ_TRY(ltree.alloc(2 * N * 32));
    SCZ_TRY(fr_pointwise(ctx, 2, pk->local_s_p, pk->sid_p, pk->alpha_beta, num.p, hl));      // :326-331
    SCZ_TRY(fr_pointwise(ctx, 2, pk->eq_r1_p, pk->ssigma_p, pk->alpha_beta, den.p, hl));     // :332-337
    SCZ_TRY(fr_pointwise(ctx, 3, num.p, den.p, nullptr, h_p.p, hl));                         // :339
    SCZ_TRY(d_acc_product_dev(ctx, h_p.p, hl, subtree.p, ltree.p));                         // :342
    const void *v1x = (const char *)subtree.p + hl * 32;                                    // skip(len/2) :344-348
    SCZ_TRY(fr_deinterleave(ctx, subtree.p, hl, vx0.p, vx1.p));                             // :349-359
    {
        const void *tabs[8] = {pk->ssigma_p, pk->sid_p, h_p.p, num.p, den.p, v1x, vx0.p, vx1.p};
        for (int k = 0; k < 8; k++) SCZ_TRY(d_commit(tabs[k], hl));                         // :363-380
        for (int k = 0; k < 5; k++) SCZ_TRY(d_open(SCZ_HP_WIRING_OPEN, tabs[k], hl, r2, n + 2, 0));   // :383-407
    }
    SCZ_TRY(early(2));   // ~3/4 of the proof's MSM work has been queued by now
    for (auto &f : later) SCZ_TRY(f());   // the gate-identity sumchecks and the sumcheck of 2.c
    SCZ_TRY(d_sum(den.p, pk->eq_r2_p, hl, r2));                                             // 2.e.1 :411-413
    SCZ_TRY(d_sum(h_p.p, den.p, hl, r2));
    SCZ_TRY(d_sum(num.p, pk->eq_r2_p, hl, r2));
    {
        // 2.e.2 (:418-478): layered zerocheck; the current tables are the first half (:419-422), then always the
        // second half of what is left (:474-477): plain pointer arithmetic, no copies
        size_t off = 0, len = hl / 2;
        for (size_t i = 1; i + s <= n; i++) {
            const void *c_v1x = (const char *)v1x + off * 32, *c_vx0 = (const char *)vx0.p + off * 32;
            const void *c_vx1 = (const char *)vx1.p + off * 32, *c_eq = (const char *)pk->eq_r2_p + off * 32;
            const void *chi = r2 + i * 32;
            size_t np = n + 2 - i;
            SCZ_TRY(d_sum(c_eq, c_v1x, len, chi));                                          // :426-435
            SCZ_TRY(d_sum(c_eq, c_vx0, len, chi));
            SCZ_TRY(d_sum(c_vx0, c_vx1, len, chi));
            SCZ_TRY(d_open(SCZ_HP_WIRING_OPEN, c_v1x, len, chi, np, 0));                    // :458-472
            SCZ_TRY(d_open(SCZ_HP_WIRING_OPEN, c_vx0, len, chi, np, 0));
            SCZ_TRY(d_open(SCZ_HP_WIRING_OPEN, c_vx1, len, chi, np, 0));
            off += len / 2;
            len /= 2;
        }
    }
    if (leader) {   // :480-511  `if let Some(leader_tree) = top`
        DevTmp *lx = nullptr;   // vx0 | vx1 of the leader tree; read by queued MSMs, so it lives in D
        SCZ_TRY(D.tmp(2 * N * 32, &lx));
        void *lx0 = lx->p, *lx1 = (char *)lx->p + N * 32;
        const void *l1x = (const char *)ltree.p + N * 32;
        SCZ_TRY(fr_deinterleave(ctx, ltree.p, N, lx0, lx1));
        const void *tabs[3] = {lx0, lx1, l1x};                                          // vx0, vx1, v1x :500-505
        for (int k = 0; k < 3; k++) {
            SCZ_TRY(o.reserve(0, 1));
            SCZ_TRY(commit_defer(ctx, D, pk->d_commitment, tabs[k], N, o.pts_at()));
            o.push(SCZ_HP_WIRING_COMMIT, 0, 1, 0);
            SCZ_TRY(o.reserve(0, s));
            SCZ_TRY(open_defer(ctx, D, pk->d_commitment, tabs[k], N, r2, o.val_at(), o.pts_at()));
            o.push(SCZ_HP_WIRING_OPEN, 0, s, 1);
        }
        const void *fs[3] = {pk->eq_leader, pk->eq_leader, lx0}, *gs[3] = {l1x, lx0, lx1};   // :507-509
        for (int k = 0; k < 3; k++) {
            SCZ_TRY(o.reserve(s + 1, 0));
            SCZ_TRY(sumcheck_product_dev(ctx, fs[k], gs[k], N, r2, o.tri_at()));
            o.push(SCZ_HP_WIRING_PROOF, s + 1, 0, 0);
        }
    }

    // ---- Open (:517-554)
    if (variant != HP_PERMCHECK) {
        void *coms_p = coms.p;
        auto copy_com = [&](void *dst, int k) {   // the commitments of step 1 exist once D has run their leader rounds
            D.then2([=]() -> int32_t {   // after the round's scatters, which deliver the commitments of step 1
                SCZ_CUDA(ctx, cudaMemcpyAsync(dst, (char *)coms_p + k * PT, PT, cudaMemcpyDeviceToDevice, st));
                return SCZ_OK;
            });
        };
        const void *tabs[3] = {pk->a_evals, pk->b_evals, pk->c_evals};
        for (int k = 0; k < 3; k++) {
            SCZ_TRY(o.reserve(0, 1));
            copy_com(o.pts_at(), k);
            SCZ_TRY(c_open(SCZ_HP_GATE_COMMIT, tabs[k], share_len, pk->challenge, 1));
        }
        const void *slc[3] = {pk->I_p, pk->S1_p, pk->S2_p};
        for (int k = 0; k < 3; k++) {
            SCZ_TRY(o.reserve(0, 1));
            copy_com(o.pts_at(), 3 + k);
            SCZ_TRY(d_open(SCZ_HP_GATE_COMMIT, slc[k], slice_len, pk->challenge, n, 1));
        }
    }
    return D.run();
}

// ---- local_hyperplonk: hyperplonk/src/hyperplonk.rs:15-160, the monolithic (single-prover) baseline the reference times
// as "Local HyperPlonk".  Same primitives, plain tables of gc = 2^n (a, b, c, input, q1, q2, eq) and 4 gc (m, ssigma, sid,
// eq_p2) entries, one SRS with levels 0 .. n+2 (new_toy, dpoly_comm.rs:107-135).
int32_t local_hyperplonk_dev(Ctx *ctx, size_t n, const scz_local_pk *pk, HpOut &o) {
    if (n < 1 || n > 26) return ctx->fail(SCZ_ERR_BAD_ARG, "local_hyperplonk: n = %zu", n);
    const size_t gc = (size_t)1 << n, g4 = gc * 4, PT = SCZ_G1_JAC_BYTES;
    Deferred D(ctx);
    DevTmp coms(ctx), tmp(ctx), num(ctx), den(ctx), h(ctx), tree(ctx), vx0(ctx), vx1(ctx);
    SCZ_TRY(coms.alloc(6 * PT));
    SCZ_TRY(tmp.alloc(gc * 32));
    SCZ_TRY(num.alloc(g4 * 32));
    SCZ_TRY(den.alloc(g4 * 32));
    SCZ_TRY(h.alloc(g4 * 32));
    SCZ_TRY(tree.alloc(2 * g4 * 32));
    SCZ_TRY(vx0.alloc(g4 * 32));
    SCZ_TRY(vx1.alloc(g4 * 32));
    const void *gate_tabs[6] = {pk->a_evals, pk->b_evals, pk->c_evals, pk->input, pk->q1, pk->q2};
    for (int k = 0; k < 6; k++) SCZ_TRY(commit_defer(ctx, D, pk->commitment, gate_tabs[k], gc, (char *)coms.p + k * PT));   // :55-62
    auto sum = [&](uint32_t kind, const void *f, const void *g, size_t len, const void *challenge) -> int32_t {
        size_t cnt = ilog2(len) + 1;
        SCZ_TRY(o.reserve(cnt, 0));
        SCZ_TRY(sumcheck_product_dev(ctx, f, g, len, challenge, o.tri_at()));
        o.push(kind, cnt, 0, 0);
        return SCZ_OK;
    };
    // gate identity (:70-96)
    SCZ_TRY(sum(SCZ_HP_GATE_PROOF, pk->eq, pk->q1, gc, pk->challenge));
    SCZ_TRY(fr_pointwise(ctx, 0, pk->a_evals, pk->b_evals, nullptr, tmp.p, gc));
    SCZ_TRY(sum(SCZ_HP_GATE_PROOF, pk->q1, tmp.p, gc, pk->challenge));
    SCZ_TRY(sum(SCZ_HP_GATE_PROOF, pk->eq, pk->q2, gc, pk->challenge));
    SCZ_TRY(sum(SCZ_HP_GATE_PROOF, pk->a_evals, pk->b_evals, gc, pk->challenge));
    SCZ_TRY(sum(SCZ_HP_GATE_PROOF, pk->q2, pk->a_evals, gc, pk->challenge));
    SCZ_TRY(fr_pointwise(ctx, 1, pk->c_evals, pk->input, nullptr, tmp.p, gc));
    SCZ_TRY(sum(SCZ_HP_GATE_PROOF, pk->eq, tmp.p, gc, pk->challenge));
    // wire identity (:98-145)
    SCZ_TRY(fr_pointwise(ctx, 2, pk->m, pk->sid, pk->alpha_beta, num.p, g4));       // :106-110  (alpha | beta)
    SCZ_TRY(fr_pointwise(ctx, 2, pk->m, pk->ssigma, pk->alpha_beta, den.p, g4));    // :111-115
    SCZ_TRY(fr_pointwise(ctx, 3, num.p, den.p, nullptr, h.p, g4));                  // :116
    SCZ_TRY(acc_product_tree(ctx, h.p, g4, tree.p));                                // :118
    SCZ_TRY(fr_deinterleave(ctx, tree.p, g4, vx0.p, vx1.p));
    const void *v1x = (const char *)tree.p + g4 * 32;
    const void *wtabs[8] = {pk->sid, pk->ssigma, h.p, num.p, den.p, vx0.p, vx1.p, v1x};   // :121-136
    for (int k = 0; k < 8; k++) {
        SCZ_TRY(o.reserve(0, 1));
        SCZ_TRY(commit_defer(ctx, D, pk->commitment, wtabs[k], g4, o.pts_at()));
        o.push(SCZ_HP_WIRING_COMMIT, 0, 1, 0);
        SCZ_TRY(o.reserve(0, n + 2));
        SCZ_TRY(open_defer(ctx, D, pk->commitment, wtabs[k], g4, pk->challengep2, o.val_at(), o.pts_at()));
        o.push(SCZ_HP_WIRING_OPEN, 0, n + 2, 1);
    }
    SCZ_TRY(sum(SCZ_HP_WIRING_PROOF, pk->eq_p2, v1x, g4, pk->challengep2));         // :138-145
    SCZ_TRY(sum(SCZ_HP_WIRING_PROOF, pk->eq_p2, vx0.p, g4, pk->challengep2));
    SCZ_TRY(sum(SCZ_HP_WIRING_PROOF, vx0.p, vx1.p, g4, pk->challengep2));
    SCZ_TRY(sum(SCZ_HP_WIRING_PROOF, pk->eq_p2, den.p, g4, pk->challengep2));
    SCZ_TRY(sum(SCZ_HP_WIRING_PROOF, h.p, den.p, g4, pk->challengep2));
    SCZ_TRY(sum(SCZ_HP_WIRING_PROOF, pk->eq_p2, num.p, g4, pk->challengep2));
    // open (:149-156)
    cudaStream_t st = ctx->stream;
    void *coms_p = coms.p;
    for (int k = 0; k < 6; k++) {
        SCZ_TRY(o.reserve(0, 1 + n));
        void *dst = o.pts_at();
        D.then2([=]() -> int32_t {
            SCZ_CUDA(ctx, cudaMemcpyAsync(dst, (char *)coms_p + k * PT, PT, cudaMemcpyDeviceToDevice, st));
            return SCZ_OK;
        });
        SCZ_TRY(open_defer(ctx, D, pk->commitment, gate_tabs[k], gc, pk->challenge, o.val_at(), o.pts_at(1)));
        o.push(SCZ_HP_GATE_COMMIT, 0, 1 + n, 1);
    }
    return D.run();
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_dhyperplonk_sizes(size_t n, size_t l, size_t n_parties, size_t *triples, size_t *points, size_t *values,
                              size_t *items) {
    if (!n || !l || !n_parties) return SCZ_ERR_BAD_ARG;
    size_t ll = ilog2(l), s = ilog2(n_parties);
    size_t n_items = 6 + 6 + (1 + 3 + 3 * n + 3) + (1 + 8 + 3) + (3 + 5 + 3 * n + 3);
    size_t per = n + 2 + ll + s + 2;   // longest entry: a c_/d_ proof on the 2^(n+2) tables
    if (items) *items = n_items;
    if (triples) *triples = n_items * per;
    if (points) *points = n_items * per;
    if (values) *values = n_items;
    return SCZ_OK;
}

static int32_t hp_entry(scz_ctx *h, size_t n, const scz_hp_pk *pk, const scz_pp *pp, void *d_triples, size_t triples_cap,
                        void *d_points, size_t points_cap, void *d_values, size_t values_cap, scz_hp_item *items,
                        size_t items_cap, size_t *n_items, HpVariant variant) {
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (!pk || !pp || !d_triples || !d_points || !d_values || !items || !n_items)
        return c->fail(SCZ_ERR_BAD_ARG, "dhyperplonk: null argument");
    const void *need[] = {pk->V, pk->a_evals, pk->b_evals, pk->c_evals, pk->I, pk->S1, pk->S2, pk->I_p, pk->S1_p, pk->S2_p,
                          pk->ssigma_p, pk->sid_p, pk->eq, pk->eq_r1_p, pk->eq_r2_p, pk->challenge, pk->challenge_r1,
                          pk->challenge_r2, pk->alpha_beta, pk->c_commitment, pk->d_commitment, pk->local_s_p, pk->local_s,
                          pk->eq_leader};
    for (const void *p : need)
        if (!p) return c->fail(SCZ_ERR_BAD_ARG, "dhyperplonk: a field of scz_hp_pk is null");
    HpOut o;
    o.ctx = c;
    o.tri = (char *)d_triples, o.pts = (char *)d_points, o.val = (char *)d_values;
    o.tri_cap = triples_cap, o.pts_cap = points_cap, o.val_cap = values_cap, o.items_cap = items_cap;
    o.items = items;
    int32_t rc = dhyperplonk_dev(c, n, pk, pp, o, variant);
    *n_items = o.items_n;
    return rc;
}

int32_t scz_local_hyperplonk_dev(scz_ctx *h, size_t n, const scz_local_pk *pk, void *d_triples, size_t triples_cap, void *d_points,
                                 size_t points_cap, void *d_values, size_t values_cap, scz_hp_item *items, size_t items_cap,
                                 size_t *n_items) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (!pk || !d_triples || !d_points || !d_values || !items || !n_items) return c->fail(SCZ_ERR_BAD_ARG, "local_hyperplonk: null argument");
    const void *need[] = {pk->m, pk->a_evals, pk->b_evals, pk->c_evals, pk->input, pk->q1, pk->q2, pk->ssigma, pk->sid, pk->eq,
                          pk->eq_p2, pk->challenge, pk->challengep2, pk->alpha_beta, pk->commitment};
    for (const void *p : need)
        if (!p) return c->fail(SCZ_ERR_BAD_ARG, "local_hyperplonk: a field of scz_local_pk is null");
    HpOut o;
    o.ctx = c;
    o.tri = (char *)d_triples, o.pts = (char *)d_points, o.val = (char *)d_values;
    o.tri_cap = triples_cap, o.pts_cap = points_cap, o.val_cap = values_cap, o.items_cap = items_cap;
    o.items = items;
    int32_t rc = local_hyperplonk_dev(c, n, pk, o);
    *n_items = o.items_n;
    return rc;
}

int32_t scz_dhyperplonk_dev(scz_ctx *h, size_t n, const scz_hp_pk *pk, const scz_pp *pp, void *d_triples, size_t triples_cap,
                            void *d_points, size_t points_cap, void *d_values, size_t values_cap, scz_hp_item *items,
                            size_t items_cap, size_t *n_items) {
    scz::DeviceGuard dg__(h);
    return hp_entry(h, n, pk, pp, d_triples, triples_cap, d_points, points_cap, d_values, values_cap, items, items_cap, n_items,
                    HP_FULL);
}
int32_t scz_dhyperplonk_data_parallel_dev(scz_ctx *h, size_t n, const scz_hp_pk *pk, const scz_pp *pp, void *d_triples,
                                          size_t triples_cap, void *d_points, size_t points_cap, void *d_values,
                                          size_t values_cap, scz_hp_item *items, size_t items_cap, size_t *n_items) {
    scz::DeviceGuard dg__(h);
    return hp_entry(h, n, pk, pp, d_triples, triples_cap, d_points, points_cap, d_values, values_cap, items, items_cap, n_items,
                    HP_DATA_PARALLEL);
}
int32_t scz_dpermcheck_dev(scz_ctx *h, size_t n, const scz_hp_pk *pk, const scz_pp *pp, void *d_triples, size_t triples_cap,
                           void *d_points, size_t points_cap, void *d_values, size_t values_cap, scz_hp_item *items,
                           size_t items_cap, size_t *n_items) {
    scz::DeviceGuard dg__(h);
    return hp_entry(h, n, pk, pp, d_triples, triples_cap, d_points, points_cap, d_values, values_cap, items, items_cap, n_items,
                    HP_PERMCHECK);
}

}   // extern "C"
