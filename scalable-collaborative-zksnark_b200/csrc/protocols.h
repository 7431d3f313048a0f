// Internal interface of the protocol layer (protocols.cu, poly.cu, dmsm.cu) for the prover schedule
// (hyperplonk.cu).  Every function is the device-side restatement of one reference function; see the
// definitions for the file:line each follows.
#pragma once
#include "ctx.h"
#include "deferred.h"
#include "pss.h"

struct scz_srs;

namespace scz {

// poly.cu
int32_t sumcheck_product_rounds(Ctx *ctx, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                                void *d_out, void *d_last);
int32_t sumcheck_rounds(Ctx *ctx, const void *d_f, size_t len, const void *d_challenge, void *d_out, void *d_last);
int32_t open_fold_rounds(Ctx *ctx, const void *d_peval, size_t len, const void *d_point, void *d_q, void *d_value);
int32_t acc_product_tree(Ctx *ctx, const void *d_x, size_t m, void *d_tree);
int32_t fr_pointwise(Ctx *c, int32_t mode, const void *d_a, const void *d_b, const void *d_k, void *d_out, size_t n);
int32_t fr_deinterleave(Ctx *c, const void *d_in, size_t n_pairs, void *d_even, void *d_odd);

// dmsm.cu
int32_t d_msm_dev(Ctx *ctx, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars,
                  const size_t *lens, size_t batch, void *d_out);

// protocols.cu
int32_t pss2ss_dev(Ctx *ctx, const scz_pp *pp, const void *d_share, void *d_out);
int32_t degree_reduce_dev(Ctx *ctx, const scz_pp *pp, const void *d_share, void *d_out);
int32_t sumcheck_product_dev(Ctx *ctx, const void *d_f, const void *d_g, size_t len, const void *d_challenge, void *d_out);
int32_t c_sumcheck_product_dev(Ctx *ctx, const scz_pp *pp, const void *d_f, const void *d_g, size_t len,
                               const void *d_challenge, void *d_out);
int32_t d_sumcheck_product_dev(Ctx *ctx, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                               void *d_out, size_t *count);
int32_t sumcheck_dev(Ctx *ctx, const void *d_f, size_t len, const void *d_challenge, void *d_out);
int32_t c_sumcheck_dev(Ctx *ctx, const scz_pp *pp, const void *d_f, size_t len, const void *d_challenge, void *d_out);
int32_t d_sumcheck_dev(Ctx *ctx, const void *d_f, size_t len, const void *d_challenge, void *d_out, size_t *count);
int32_t d_acc_product_dev(Ctx *ctx, const void *d_x, size_t m, void *d_subtree, void *d_leader_tree);
int32_t commit_dev(Ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, void *d_out);
int32_t c_commit_dev(Ctx *ctx, const scz_srs *srs, const scz_pp *pp, const void *const *d_pevals, const size_t *lens,
                     size_t batch, void *d_out);
int32_t d_commit_dev(Ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, void *d_out);
int32_t open_dev(Ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point, void *d_value,
                 void *d_proofs);
int32_t c_open_dev(Ctx *ctx, const scz_srs *srs, const scz_pp *pp, const void *d_peval, size_t len, const void *d_point,
                   void *d_value, void *d_proofs);
int32_t d_open_dev(Ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point, size_t npoint,
                   void *d_value, void *d_proofs, size_t *count);


// deferred variants (deferred.h): MSMs are queued on D, leader rounds become continuations
int32_t d_msm_defer(Ctx *ctx, Deferred &D, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars,
                    const size_t *lens, size_t batch, void *d_out, const uint32_t *pre_c = nullptr);
int32_t commit_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const void *d_peval, size_t len, void *d_out);
int32_t c_commit_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const scz_pp *pp, const void *const *d_pevals,
                       const size_t *lens, size_t batch, void *d_out);
int32_t d_commit_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const void *d_peval, size_t len, void *d_out);
int32_t open_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point,
                   void *d_value, void *d_proofs);
int32_t c_open_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const scz_pp *pp, const void *d_peval, size_t len,
                     const void *d_point, void *d_value, void *d_proofs);
int32_t d_open_defer(Ctx *ctx, Deferred &D, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point,
                     size_t npoint, void *d_value, void *d_proofs, size_t *count);

}   // namespace scz
