// d_msm: dist-primitive/src/dmsm.rs:9-43.
//   c_shares[k] = G::msm(bases[k], scalars[k])                       (:19-24)  -> one batched Pippenger sequence
//   leader_compute_element(c_shares, f)                              (:29-40)
//     f: transpose; per k: unpack2 -> sum of the l secrets -> [sum; l] -> pack_from_public; transpose
// The gather / scatter payloads stay on the device in Jacobian form (144 B per
// point); the byte counters use the reference's wire size, Vec<G1> compressed =
// 8 + 48 * batch (ark-serialize, serializing_net.rs:17).
#include <vector>

#include "deferred.h"
#include "g1.cuh"
#include "msm.h"
#include "net.h"
#include "pss.h"

namespace scz {

// The leader closure alone (dmsm.rs:31-38) on an already gathered buffer: recv is party-major
// [party][k] (n x batch Jacobian points), send receives the same layout.  unpack2 -> sum of the l
// secrets -> replicate -> pack_from_public is ONE fixed rank-one map (u, p built at scz_pp_new): n scalar
// multiplications into S_k, n more out of it, per batch entry (pss.cu, k_pss_dmsm_multi).
int32_t d_msm_leader(Ctx *ctx, const scz_pp *pp, const void *d_recv, size_t batch, void *d_send) {
    ProfScope ps(ctx, SCZ_K_PSS);
    struct { const void *in; void *out; uint32_t batch, cta_base; } job = {d_recv, d_send, (uint32_t)batch, 0};
    return pss_dmsm_multi(ctx, pp, &job, 1);
}

// Queues the local MSMs (dmsm.rs:19-24) on `D` and registers the leader round (:29-40) as their continuation.
// pre_c (optional): per entry, the window of a fixed-base table passed as `d_bases[k]` (srs.cu); 0 = plain bases
int32_t d_msm_defer(Ctx *ctx, Deferred &D, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars,
                    const size_t *lens, size_t batch, void *d_out, const uint32_t *pre_c) {
    if (!pp) return ctx->fail(SCZ_ERR_BAD_ARG, "d_msm: null pp");
    if (batch == 0) return SCZ_OK;
    Net *net = ctx->net;
    const size_t N = net->n_parties, PT = SCZ_G1_JAC_BYTES;
    if (N != pp->n) return ctx->fail(SCZ_ERR_BAD_ARG, "d_msm: %zu parties but pp.n = %zu", N, pp->n);
    DevTmp *c_shares = nullptr;
    SCZ_TRY(D.tmp(batch * PT, &c_shares));
    for (size_t k = 0; k < batch; k++)
        SCZ_TRY(D.add_msm(d_bases[k], d_scalars[k], lens[k], (char *)c_shares->p + k * PT, pre_c ? pre_c[k] : 0));
    Deferred *Dp = &D;
    D.then([=]() -> int32_t {
        const size_t wire = 8 + 48 * batch;
        DevTmp *recv = nullptr, *send = nullptr;
        if (net->is_leader()) {
            SCZ_TRY(Dp->tmp(N * batch * PT, &recv));
            SCZ_TRY(Dp->tmp(N * batch * PT, &send));
        }
        Dp->gather(c_shares->p, recv ? recv->p : nullptr, batch * PT, wire);
        // recv is party-major [j][k]: vector k is the stride-`batch` column.  The closure (dmsm.rs:31-38) is queued:
        // all d_msm calls of a round share one launch, and one gather / one scatter (deferred.h)
        if (net->is_leader()) Dp->add_pss(pp, recv->p, (uint32_t)batch, send->p);
        Dp->scatter(send ? send->p : nullptr, d_out, batch * PT, wire);
        return SCZ_OK;
    });
    return SCZ_OK;
}

int32_t d_msm_dev(Ctx *ctx, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars,
                  const size_t *lens, size_t batch, void *d_out) {
    Deferred D(ctx);
    SCZ_TRY(d_msm_defer(ctx, D, pp, d_bases, d_scalars, lens, batch, d_out, nullptr));
    return D.run();
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_d_msm_dev(scz_ctx *h, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars,
                      const size_t *lens, size_t batch, void *d_out) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (batch && (!d_bases || !d_scalars || !lens || !d_out)) return h->c.fail(SCZ_ERR_BAD_ARG, "d_msm: null argument");
    return d_msm_dev(&h->c, pp, d_bases, d_scalars, lens, batch, d_out);
}

int32_t scz_d_msm_leader_dev(scz_ctx *h, const scz_pp *pp, const void *d_gathered, size_t batch, void *d_to_scatter) {
    scz::DeviceGuard dg__(h);
    if (!h || !pp) return SCZ_ERR_BAD_ARG;
    if (batch && (!d_gathered || !d_to_scatter)) return h->c.fail(SCZ_ERR_BAD_ARG, "d_msm_leader: null argument");
    if (!batch) return SCZ_OK;
    return d_msm_leader(&h->c, pp, d_gathered, batch, d_to_scatter);
}

int32_t scz_d_msm(scz_ctx *h, const scz_pp *pp, const void *const *bases, const size_t *bases_lens,
                  const void *const *scalars, const size_t *scalars_lens, size_t batch, void *out_jac) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (batch && (!bases || !scalars || !bases_lens || !scalars_lens || !out_jac))
        return c->fail(SCZ_ERR_BAD_ARG, "d_msm: null argument");
    size_t tot = 0;
    for (size_t k = 0; k < batch; k++) {
        if (bases_lens[k] != scalars_lens[k])   // G::msm(..).unwrap() panics in the reference (dmsm.rs:23)
            return c->fail(SCZ_ERR_LEN_MISMATCH, "d_msm: entry %zu has %zu bases vs %zu scalars", k, bases_lens[k],
                           scalars_lens[k]);
        tot += bases_lens[k];
    }
    DevTmp d_b(c), d_s(c), d_o(c);
    SCZ_TRY(d_b.alloc(tot * SCZ_G1_AFFINE_BYTES));
    SCZ_TRY(d_s.alloc(tot * SCZ_FR_BYTES));
    SCZ_TRY(d_o.alloc(batch * SCZ_G1_JAC_BYTES));
    std::vector<const void *> bp(batch), sp(batch);
    size_t off = 0;
    for (size_t k = 0; k < batch; k++) {
        size_t n = bases_lens[k];
        bp[k] = d_b.as<char>() + off * SCZ_G1_AFFINE_BYTES;
        sp[k] = d_s.as<char>() + off * SCZ_FR_BYTES;
        if (n) {
            SCZ_CUDA(c, cudaMemcpyAsync((void *)bp[k], bases[k], n * SCZ_G1_AFFINE_BYTES, cudaMemcpyHostToDevice, c->stream));
            SCZ_CUDA(c, cudaMemcpyAsync((void *)sp[k], scalars[k], n * SCZ_FR_BYTES, cudaMemcpyHostToDevice, c->stream));
        }
        off += n;
    }
    SCZ_TRY(d_msm_dev(c, pp, bp.data(), sp.data(), bases_lens, batch, d_o.p));
    SCZ_CUDA(c, cudaMemcpyAsync(out_jac, d_o.p, batch * SCZ_G1_JAC_BYTES, cudaMemcpyDeviceToHost, c->stream));
    SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    return SCZ_OK;
}

}   // extern "C"
