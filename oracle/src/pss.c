/*
 * ORACLE (test infrastructure): packed secret sharing.
 * Follows secret-sharing/src/pss.rs:38-171 (PackedSharingParams::{new,
 * pack_from_public, pack_single, unpack, unpack2}) over Fr and over G1, the way
 * the reference runs ark-poly 0.4.2 Radix2EvaluationDomain FFTs on any
 * DomainCoeff.  An FFT over a size-n (coset) domain is the DFT
 *     out[j] = sum_i v[i] * (offset * w^j)^i ,
 * and ark-poly's `fft_in_place` / `ifft_in_place` first RESIZE the vector to
 * the domain size (zero-pad, or truncate when longer) -- that truncation is
 * what makes `unpack` ignore high coefficients (pss.rs:145) and gives
 * `pack_single` (pss.rs:103-113) its "pack twice" behaviour.  Sizes here are
 * tiny (n = 8l), so the DFT is evaluated directly; results are the same field
 * elements an FFT produces.
 */
#include <stdlib.h>
#include "oracle.h"

typedef union { fr_t f; g1j_t g; } elem_t;

static size_t esize(int kind) { return kind == 0 ? sizeof(fr_t) : sizeof(g1j_t); }
static void e_zero(int kind, elem_t *r) {
    if (kind == 0) fr_set_zero(&r->f);
    else g1j_set_inf(&r->g);
}
static void e_add(int kind, elem_t *r, const elem_t *a, const elem_t *b) {
    if (kind == 0) fr_add(&r->f, &a->f, &b->f);
    else g1j_add(&r->g, &a->g, &b->g);
}
static void e_mul(int kind, elem_t *r, const elem_t *a, const fr_t *k) {
    if (kind == 0) fr_mul(&r->f, &a->f, k);
    else g1j_mul_fr(&r->g, &a->g, k);
}
static void e_load(int kind, elem_t *r, const void *arr, size_t i) {
    memcpy(r, (const char *)arr + i * esize(kind), esize(kind));
}
static void e_store(int kind, void *arr, size_t i, const elem_t *v) {
    memcpy((char *)arr + i * esize(kind), v, esize(kind));
}

/* F::get_root_of_unity(n): TWO_ADIC_ROOT^(2^(32 - log n)), TWO_ADIC_ROOT = 7^((r-1)/2^32) */
static void root_of_unity(fr_t *w, size_t n) {
    fr_t g;
    fr_from_u64(&g, 7);
    uint64_t e[4];
    fr_t rm1 = fr_MOD;
    rm1.l[0] -= 1;
    /* (r-1) >> 32 */
    for (int i = 0; i < 4; i++) e[i] = (rm1.l[i] >> 32) | (i < 3 ? rm1.l[i + 1] << 32 : 0);
    fr_pow(w, &g, e, 4);
    size_t lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    for (size_t i = 0; i < 32 - lg; i++) fr_sqr(w, w);
}

void orc_pp_new(orc_pp_t *pp, size_t l) {
    orc_init();
    pp->l = l;
    pp->n = 8 * l;
    pp->t = l - 1;
    root_of_unity(&pp->share_gen, pp->n);
    root_of_unity(&pp->secret_gen, 2 * l);      /* l + t + 1 */
    root_of_unity(&pp->secret2_gen, 4 * l);
    fr_from_u64(&pp->coset, 7);
}

/* v (len entries) -> evaluations on {offset * gen^j}, j < size; v is resized to size first */
static void dft(int kind, const fr_t *gen, const fr_t *offset, size_t size, const elem_t *v, size_t len,
                elem_t *out) {
    fr_t xj = *offset;             /* offset * gen^j */
    for (size_t j = 0; j < size; j++) {
        elem_t acc, t;
        e_zero(kind, &acc);
        fr_t pw = fr_R1;
        for (size_t i = 0; i < size; i++) {
            if (i < len) {
                e_mul(kind, &t, &v[i], &pw);
                e_add(kind, &acc, &acc, &t);
            }
            fr_mul(&pw, &pw, &xj);
        }
        out[j] = acc;
        fr_mul(&xj, &xj, gen);
    }
}
/* inverse of dft: evaluations (resized to size) -> coefficients */
static void idft(int kind, const fr_t *gen, const fr_t *offset, size_t size, const elem_t *v, size_t len,
                 elem_t *out) {
    fr_t gen_inv, off_inv, n_inv, nf;
    fr_inv(&gen_inv, gen);
    fr_inv(&off_inv, offset);
    fr_from_u64(&nf, size);
    fr_inv(&n_inv, &nf);
    fr_t wi = fr_R1;               /* gen^-i */
    fr_t scale = n_inv;            /* n^-1 * offset^-i */
    for (size_t i = 0; i < size; i++) {
        elem_t acc, t;
        e_zero(kind, &acc);
        fr_t pw = fr_R1;           /* gen^(-i*j) */
        for (size_t j = 0; j < size; j++) {
            if (j < len) {
                e_mul(kind, &t, &v[j], &pw);
                e_add(kind, &acc, &acc, &t);
            }
            fr_mul(&pw, &pw, &wi);
        }
        e_mul(kind, &out[i], &acc, &scale);
        fr_mul(&wi, &wi, &gen_inv);
        fr_mul(&scale, &scale, &off_inv);
    }
}

static elem_t *load_all(int kind, const void *in, size_t len) {
    elem_t *v = malloc((len ? len : 1) * sizeof *v);
    for (size_t i = 0; i < len; i++) e_load(kind, &v[i], in, i);
    return v;
}

/* pss.rs:93-99: secret.ifft_in_place ; share.fft_in_place */
static void pack_in_place(const orc_pp_t *pp, int kind, elem_t *v, size_t len, elem_t *out /* n */) {
    size_t s = 2 * pp->l;
    elem_t *coef = malloc(s * sizeof *coef);
    idft(kind, &pp->secret_gen, &pp->coset, s, v, len < s ? len : s, coef);
    dft(kind, &pp->share_gen, &fr_R1, pp->n, coef, s, out);
    free(coef);
}
void orc_pack_from_public(const orc_pp_t *pp, int kind, const void *in, size_t len_in, void *out) {
    elem_t *v = load_all(kind, in, len_in);
    elem_t *o = malloc(pp->n * sizeof *o);
    pack_in_place(pp, kind, v, len_in, o);
    for (size_t i = 0; i < pp->n; i++) e_store(kind, out, i, &o[i]);
    free(v);
    free(o);
}
/* pss.rs:103-113: pack [s], then pack the resulting n shares AGAIN (the
 * second ifft truncates them to the first 2l). */
void orc_pack_single(const orc_pp_t *pp, int kind, const void *in, void *out) {
    elem_t *v = load_all(kind, in, 1);
    elem_t *o1 = malloc(pp->n * sizeof *o1), *o2 = malloc(pp->n * sizeof *o2);
    pack_in_place(pp, kind, v, 1, o1);
    pack_in_place(pp, kind, o1, pp->n, o2);
    for (size_t i = 0; i < pp->n; i++) e_store(kind, out, i, &o2[i]);
    free(v);
    free(o1);
    free(o2);
}
/* pss.rs:132-149 */
void orc_unpack(const orc_pp_t *pp, int kind, const void *in, void *out) {
    size_t s = 2 * pp->l;
    elem_t *v = load_all(kind, in, pp->n);
    elem_t *coef = malloc(pp->n * sizeof *coef), *ev = malloc(s * sizeof *ev);
    idft(kind, &pp->share_gen, &fr_R1, pp->n, v, pp->n, coef);
    dft(kind, &pp->secret_gen, &pp->coset, s, coef, s, ev);      /* truncated to 2l coefficients */
    for (size_t i = 0; i < pp->l; i++) e_store(kind, out, i, &ev[i]);
    free(v);
    free(coef);
    free(ev);
}
/* pss.rs:153-171 */
void orc_unpack2(const orc_pp_t *pp, int kind, const void *in, void *out) {
    size_t s = 4 * pp->l;
    elem_t *v = load_all(kind, in, pp->n);
    elem_t *coef = malloc(pp->n * sizeof *coef), *ev = malloc(s * sizeof *ev);
    idft(kind, &pp->share_gen, &fr_R1, pp->n, v, pp->n, coef);
    dft(kind, &pp->secret2_gen, &pp->coset, s, coef, s, ev);     /* truncated to 4l coefficients */
    for (size_t i = 0; i < pp->l; i++) e_store(kind, out, i, &ev[2 * i]);   /* step_by(2) over [0,2l) */
    free(v);
    free(coef);
    free(ev);
}
