/*
 * ORACLE -- CPU restatement of the dist-primitive hot path of
 * LBruyne/Scalable-Collaborative-zkSNARK.  TEST INFRASTRUCTURE ONLY: nothing
 * in the product library (scalable-collaborative-zksnark_b200/) may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the timed CPU arm.
 *
 * PARITY UNPINNED at the arkworks byte boundary: the reference's arithmetic
 * lives in ark-ff/ark-ec/ark-poly 0.4.2 (Cargo.lock:95-223), which are not in
 * /root/reference, no Rust toolchain exists here and the reference holds no
 * golden vectors.  What pins this file instead is listed in DESIGN.md
 * ("oracle pinning") and exercised by tests/test_oracle_*.py.
 */
#ifndef SCZ_ORACLE_H
#define SCZ_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#define FP(x) fr_##x
#define NL 4
#include "fp_tmpl.h"
#undef FP
#undef NL
#define FP(x) fq_##x
#define NL 6
#include "fp_tmpl.h"
#undef FP
#undef NL

/* ark-ec short_weierstrass::Affine{x,y,infinity} / Projective{x,y,z} (Jacobian) */
typedef struct { fq_t x, y; uint32_t inf; uint32_t pad_; } g1a_t;   /* 104 B */
typedef struct { fq_t x, y, z; } g1j_t;                             /* 144 B */

void orc_init(void);
/* tiny pthread work-sharing loop (no OpenMP runtime in this image) */
void orc_parallel_for(int threads, size_t count, void (*body)(void *ctx, size_t i), void *ctx);
int  orc_hw_threads(void);

/* ---- G1 (g1.c) ---- */
extern g1a_t G1_GEN;
void g1j_set_inf(g1j_t *r);
int  g1j_is_inf(const g1j_t *a);
void g1j_from_affine(g1j_t *r, const g1a_t *a);
void g1j_to_affine(g1a_t *r, const g1j_t *a);
void g1j_double(g1j_t *r, const g1j_t *a);
void g1j_add(g1j_t *r, const g1j_t *a, const g1j_t *b);
void g1j_add_mixed(g1j_t *r, const g1j_t *a, const g1a_t *b);
void g1j_neg(g1j_t *r, const g1j_t *a);
int  g1j_eq(const g1j_t *a, const g1j_t *b);
int  g1a_on_curve(const g1a_t *a);
/* k = Montgomery-form Fr scalar */
void g1j_mul_fr(g1j_t *r, const g1j_t *a, const fr_t *k);
/* k = canonical little-endian limbs */
void g1j_mul_bits(g1j_t *r, const g1j_t *a, const uint64_t *k, int limbs);
void g1_msm_naive(g1j_t *r, const g1a_t *bases, const fr_t *scalars, size_t n);
/* ark-ec 0.4.2 VariableBaseMSM::msm restated (signed-digit Pippenger) */
void g1_msm_ark(g1j_t *r, const g1a_t *bases, const fr_t *scalars, size_t n);
void g1_msm_ark_mt(g1j_t *r, const g1a_t *bases, const fr_t *scalars, size_t n, int threads);
void orc_set_msm_threads(int threads);

/* ---- PSS (pss.c): secret-sharing/src/pss.rs ---- */
typedef struct {
    size_t t, l, n;            /* pss.rs:38-41 */
    fr_t share_gen;            /* omega_n, offset 1 */
    fr_t secret_gen;           /* omega_{2l}, coset offset GENERATOR */
    fr_t secret2_gen;          /* omega_{4l}, coset offset GENERATOR */
    fr_t coset;                /* F::GENERATOR = 7 */
} orc_pp_t;
void orc_pp_new(orc_pp_t *pp, size_t l);
/* element kind: 0 = Fr (fr_t), 1 = G1 Jacobian (g1j_t).  `in` has len_in
 * elements; out receives n (pack*) or l (unpack*) elements. */
void orc_pack_from_public(const orc_pp_t *pp, int kind, const void *in, size_t len_in, void *out);
void orc_pack_single(const orc_pp_t *pp, int kind, const void *in, void *out);
void orc_unpack(const orc_pp_t *pp, int kind, const void *in, void *out);
void orc_unpack2(const orc_pp_t *pp, int kind, const void *in, void *out);

/* ---- "network" modes (net.c): dist-primitive/src/utils/serializing_net.rs ---- */
enum { ORC_LEADER_SIM = 0, ORC_PARTIES = 1 };

/* ---- dmsm.c: dist-primitive/src/dmsm.rs:9-43 ----
 * mode LEADER_SIM: one party (the leader), inputs indexed [k]; out[k].
 * mode PARTIES:    inputs indexed [party][k]; out[party][k].
 * bases/scalars are arrays of pointers (P*batch entries), lens[k] per batch entry. */
void orc_d_msm(const orc_pp_t *pp, int mode, size_t batch, const size_t *lens,
               const g1a_t *const *bases, const fr_t *const *scalars, g1j_t *out, int use_ark_msm);

/* ---- sumcheck.c: dist-primitive/src/dsumcheck.rs ---- */
typedef struct { fr_t a, b, c; } fr3_t;
typedef struct { fr_t a, b; } fr2_t;
/* single-MLE variants (dsumcheck.rs:6-26, 92-146, 287-357); same conventions as the product ones below */
size_t orc_sumcheck(const fr_t *f, size_t len, const fr_t *challenge, fr2_t *out);
size_t orc_c_sumcheck(const orc_pp_t *pp, int mode, const fr_t *const *f, size_t len, const fr_t *challenge, fr2_t *out);
size_t orc_d_sumcheck(int mode, size_t nparties, const fr_t *const *f, size_t len, const fr_t *challenge, fr2_t *out);
/* returns number of triples written (n+1) */
size_t orc_sumcheck_product(const fr_t *f, const fr_t *g, size_t len, const fr_t *challenge, fr3_t *out);
/* c_sumcheck_product, one party's view in LEADER_SIM; in PARTIES f/g are [party] pointers, out [party][n+logl+1] */
size_t orc_c_sumcheck_product(const orc_pp_t *pp, int mode, const fr_t *const *f, const fr_t *const *g,
                              size_t len, const fr_t *challenge, fr3_t *out);
/* d_sumcheck_product: leader output only (n+s triples); nparties is N in both modes */
size_t orc_d_sumcheck_product(int mode, size_t nparties, const fr_t *const *f, const fr_t *const *g,
                              size_t len, const fr_t *challenge, fr3_t *out);
void orc_pss2ss(const orc_pp_t *pp, int mode, const fr_t *share_per_party, fr_t *out);
void orc_degree_reduce(const orc_pp_t *pp, int mode, const fr_t *share_per_party, fr_t *out);
void orc_fix_variable(const fr_t *evals, size_t len, const fr_t *points, size_t npoints, fr_t *out);

/* ---- accprod.c: dist-primitive/src/dacc_product.rs ---- */
void orc_sub_index(size_t i, size_t *x0, size_t *x1);
/* tree has 2*len entries */
void orc_acc_product_tree(const fr_t *x, size_t len, fr_t *tree);
/* d_acc_product: subtree per party (2*len each), leader_tree (2*N) */
void orc_d_acc_product(int mode, size_t nparties, const fr_t *const *inputs, size_t len,
                       fr_t *const *subtrees, fr_t *leader_tree);

/* ---- pst.c: dist-primitive/src/dpoly_comm.rs ---- */
typedef struct {
    size_t levels;             /* powers_of_g has levels entries: level i holds 2^i (or max(1,2^i/l)) points */
    g1a_t **powers_of_g;
    size_t *level_len;
} orc_srs_t;
/* PolynomialCommitmentCub::new (dpoly_comm.rs:37-67) with trapdoor s[0..n) , then mature() */
void orc_srs_new(orc_srs_t *srs, const g1j_t *g, const fr_t *s, size_t n);
/* wrap caller-provided bases (new_single / new_random shapes) */
void orc_srs_from_levels(orc_srs_t *srs, size_t levels, g1a_t **levels_ptr, const size_t *level_len);
void orc_srs_free(orc_srs_t *srs);
void orc_commit(const orc_srs_t *srs, const fr_t *peval, size_t len, g1j_t *out, int use_ark_msm);
/* open: value + n proofs */
void orc_open(const orc_srs_t *srs, const fr_t *peval, size_t len, const fr_t *point, fr_t *value,
              g1j_t *proofs, int use_ark_msm);
void orc_c_commit(const orc_srs_t *const *srs_per_party, const orc_pp_t *pp, int mode, size_t batch,
                  const size_t *lens, const fr_t *const *pevals, g1j_t *out, int use_ark_msm);
size_t orc_c_open(const orc_srs_t *const *srs_per_party, const orc_pp_t *pp, int mode,
                  const fr_t *const *peval, size_t len, const fr_t *point, fr_t *value, g1j_t *proofs,
                  int use_ark_msm);
void orc_d_commit(const orc_srs_t *const *srs_per_party, int mode, size_t nparties,
                  const fr_t *const *peval, size_t len, g1j_t *out, int use_ark_msm);
/* leader's answer: value + (logN + n) proofs; returns proof count */
size_t orc_d_open(const orc_srs_t *const *srs_per_party, int mode, size_t nparties,
                  const fr_t *const *peval, size_t len, const fr_t *point, size_t npoint, fr_t *value,
                  g1j_t *proofs, int use_ark_msm);

/* ---- misc helpers exported for the Python side ---- */
void orc_fr_mul_vec(const fr_t *a, const fr_t *b, fr_t *r, size_t n);
void orc_fr_add_vec(const fr_t *a, const fr_t *b, fr_t *r, size_t n);
void orc_fr_sub_vec(const fr_t *a, const fr_t *b, fr_t *r, size_t n);
void orc_fr_inv_vec(const fr_t *a, fr_t *r, size_t n);
void orc_fr_to_canon_vec(const fr_t *a, fr_t *r, size_t n);
void orc_fr_from_canon_vec(const fr_t *a, fr_t *r, size_t n);
void orc_fq_mul_vec(const fq_t *a, const fq_t *b, fq_t *r, size_t n);
void orc_fq_add_vec(const fq_t *a, const fq_t *b, fq_t *r, size_t n);
void orc_fq_sub_vec(const fq_t *a, const fq_t *b, fq_t *r, size_t n);
void orc_fq_to_canon_vec(const fq_t *a, fq_t *r, size_t n);
void orc_fq_from_canon_vec(const fq_t *a, fq_t *r, size_t n);
/* bases[i] = k[i] * G1_GEN as affine (k Montgomery Fr) */
void orc_g1_gen_mul_vec(const fr_t *k, g1a_t *out, size_t n);
void orc_g1j_to_affine_vec(const g1j_t *in, g1a_t *out, size_t n);
void orc_g1_add_mixed_vec(const g1j_t *a, const g1a_t *b, g1j_t *r, size_t n);
void orc_g1_add_vec(const g1j_t *a, const g1j_t *b, g1j_t *r, size_t n);
void orc_g1_double_vec(const g1j_t *a, g1j_t *r, size_t n);
#endif
