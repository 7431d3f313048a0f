/*
 * ORACLE (test infrastructure): BLS12-381 G1 group law and MSM.
 *
 * Restates what the reference calls in ark-ec 0.4.2 (Cargo.lock:119-120):
 *   - short_weierstrass::Projective (Jacobian) add / mixed add / double
 *     (EFD add-2007-bl, madd-2007-bl, dbl-2009-l for a = 0), used by every
 *     `G::msm`, `+=`, `.sum()` and `* scalar` in dist-primitive;
 *   - VariableBaseMSM::msm (call sites dist-primitive/src/dmsm.rs:23,
 *     dpoly_comm.rs:242,274,457): signed-digit windowed Pippenger with
 *     c = 3 if n < 32 else floor(log2(n)*69/100)+2, ceil(255/c) windows,
 *     2^c buckets per window, running-sum reduction, Horner combine.
 * The group element an MSM returns is algorithm-independent, so g1_msm_naive
 * (double-and-add) is the independent check and g1_msm_ark is the CPU arm that
 * bench.py times.
 */
#include <stdlib.h>
#include "oracle.h"

g1a_t G1_GEN;
static fq_t B_COEFF;   /* 4 */

static const uint64_t GEN_X[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                                  0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
static const uint64_t GEN_Y[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                                  0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};

void g1_init_(void) {
    fq_t t;
    memcpy(t.l, GEN_X, sizeof t.l);
    fq_from_canon(&G1_GEN.x, &t);
    memcpy(t.l, GEN_Y, sizeof t.l);
    fq_from_canon(&G1_GEN.y, &t);
    G1_GEN.inf = 0;
    G1_GEN.pad_ = 0;
    fq_from_u64(&B_COEFF, 4);
}

void g1j_set_inf(g1j_t *r) {
    fq_set_one(&r->x);
    fq_set_one(&r->y);
    fq_set_zero(&r->z);
}
int g1j_is_inf(const g1j_t *a) { return fq_is_zero(&a->z); }
void g1j_from_affine(g1j_t *r, const g1a_t *a) {
    if (a->inf) { g1j_set_inf(r); return; }
    r->x = a->x;
    r->y = a->y;
    fq_set_one(&r->z);
}
void g1j_to_affine(g1a_t *r, const g1j_t *a) {
    memset(r, 0, sizeof *r);
    if (g1j_is_inf(a)) { r->inf = 1; return; }
    fq_t zi, zi2, zi3;
    fq_inv(&zi, &a->z);
    fq_sqr(&zi2, &zi);
    fq_mul(&zi3, &zi2, &zi);
    fq_mul(&r->x, &a->x, &zi2);
    fq_mul(&r->y, &a->y, &zi3);
}
int g1a_on_curve(const g1a_t *a) {
    if (a->inf) return 1;
    fq_t l, r;
    fq_sqr(&l, &a->y);
    fq_sqr(&r, &a->x);
    fq_mul(&r, &r, &a->x);
    fq_add(&r, &r, &B_COEFF);
    return fq_eq(&l, &r);
}
void g1j_neg(g1j_t *r, const g1j_t *a) {
    r->x = a->x;
    fq_neg(&r->y, &a->y);
    r->z = a->z;
}
/* dbl-2009-l */
void g1j_double(g1j_t *r, const g1j_t *p) {
    if (g1j_is_inf(p)) { *r = *p; return; }
    fq_t A, B, C, D, E, F, t;
    fq_sqr(&A, &p->x);
    fq_sqr(&B, &p->y);
    fq_sqr(&C, &B);
    fq_add(&t, &p->x, &B);
    fq_sqr(&t, &t);
    fq_sub(&t, &t, &A);
    fq_sub(&t, &t, &C);
    fq_dbl(&D, &t);
    fq_dbl(&E, &A);
    fq_add(&E, &E, &A);
    fq_sqr(&F, &E);
    fq_t z3;
    fq_mul(&z3, &p->y, &p->z);
    fq_dbl(&z3, &z3);
    fq_dbl(&t, &D);
    fq_sub(&r->x, &F, &t);
    fq_sub(&t, &D, &r->x);
    fq_mul(&t, &E, &t);
    fq_dbl(&C, &C);
    fq_dbl(&C, &C);
    fq_dbl(&C, &C);
    fq_sub(&r->y, &t, &C);
    r->z = z3;
}
/* madd-2007-bl */
void g1j_add_mixed(g1j_t *r, const g1j_t *p, const g1a_t *q) {
    if (q->inf) { *r = *p; return; }
    if (g1j_is_inf(p)) { g1j_from_affine(r, q); return; }
    fq_t z1z1, u2, s2, h, hh, i, j, rr, v, t;
    fq_sqr(&z1z1, &p->z);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s2, &q->y, &p->z);
    fq_mul(&s2, &s2, &z1z1);
    if (fq_eq(&u2, &p->x)) {
        if (fq_eq(&s2, &p->y)) { g1j_double(r, p); return; }
        g1j_set_inf(r);
        return;
    }
    fq_sub(&h, &u2, &p->x);
    fq_sqr(&hh, &h);
    fq_dbl(&i, &hh);
    fq_dbl(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&rr, &s2, &p->y);
    fq_dbl(&rr, &rr);
    fq_mul(&v, &p->x, &i);
    g1j_t o;
    fq_sqr(&o.x, &rr);
    fq_sub(&o.x, &o.x, &j);
    fq_sub(&o.x, &o.x, &v);
    fq_sub(&o.x, &o.x, &v);
    fq_sub(&t, &v, &o.x);
    fq_mul(&t, &rr, &t);
    fq_mul(&j, &p->y, &j);
    fq_dbl(&j, &j);
    fq_sub(&o.y, &t, &j);
    fq_add(&t, &p->z, &h);
    fq_sqr(&t, &t);
    fq_sub(&t, &t, &z1z1);
    fq_sub(&o.z, &t, &hh);
    *r = o;
}
/* add-2007-bl */
void g1j_add(g1j_t *r, const g1j_t *p, const g1j_t *q) {
    if (g1j_is_inf(p)) { *r = *q; return; }
    if (g1j_is_inf(q)) { *r = *p; return; }
    fq_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
    fq_sqr(&z1z1, &p->z);
    fq_sqr(&z2z2, &q->z);
    fq_mul(&u1, &p->x, &z2z2);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s1, &p->y, &q->z);
    fq_mul(&s1, &s1, &z2z2);
    fq_mul(&s2, &q->y, &p->z);
    fq_mul(&s2, &s2, &z1z1);
    if (fq_eq(&u1, &u2)) {
        if (fq_eq(&s1, &s2)) { g1j_double(r, p); return; }
        g1j_set_inf(r);
        return;
    }
    fq_sub(&h, &u2, &u1);
    fq_dbl(&i, &h);
    fq_sqr(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&rr, &s2, &s1);
    fq_dbl(&rr, &rr);
    fq_mul(&v, &u1, &i);
    g1j_t o;
    fq_sqr(&o.x, &rr);
    fq_sub(&o.x, &o.x, &j);
    fq_sub(&o.x, &o.x, &v);
    fq_sub(&o.x, &o.x, &v);
    fq_sub(&t, &v, &o.x);
    fq_mul(&t, &rr, &t);
    fq_mul(&s1, &s1, &j);
    fq_dbl(&s1, &s1);
    fq_sub(&o.y, &t, &s1);
    fq_add(&t, &p->z, &q->z);
    fq_sqr(&t, &t);
    fq_sub(&t, &t, &z1z1);
    fq_sub(&t, &t, &z2z2);
    fq_mul(&o.z, &t, &h);
    *r = o;
}
/* ark-ec Projective PartialEq: cross-multiplied coordinates */
int g1j_eq(const g1j_t *a, const g1j_t *b) {
    int ia = g1j_is_inf(a), ib = g1j_is_inf(b);
    if (ia || ib) return ia && ib;
    fq_t z1z1, z2z2, l, r;
    fq_sqr(&z1z1, &a->z);
    fq_sqr(&z2z2, &b->z);
    fq_mul(&l, &a->x, &z2z2);
    fq_mul(&r, &b->x, &z1z1);
    if (!fq_eq(&l, &r)) return 0;
    fq_mul(&z1z1, &z1z1, &a->z);
    fq_mul(&z2z2, &z2z2, &b->z);
    fq_mul(&l, &a->y, &z2z2);
    fq_mul(&r, &b->y, &z1z1);
    return fq_eq(&l, &r);
}
void g1j_mul_bits(g1j_t *r, const g1j_t *a, const uint64_t *k, int limbs) {
    g1j_t acc;
    g1j_set_inf(&acc);
    for (int i = limbs * 64 - 1; i >= 0; i--) {
        g1j_double(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) g1j_add(&acc, &acc, a);
    }
    *r = acc;
}
void g1j_mul_fr(g1j_t *r, const g1j_t *a, const fr_t *k) {
    fr_t c;
    fr_to_canon(&c, k);
    g1j_mul_bits(r, a, c.l, 4);
}
void g1_msm_naive(g1j_t *r, const g1a_t *bases, const fr_t *scalars, size_t n) {
    g1j_t acc, t, b;
    g1j_set_inf(&acc);
    for (size_t i = 0; i < n; i++) {
        g1j_from_affine(&b, &bases[i]);
        g1j_mul_fr(&t, &b, &scalars[i]);
        g1j_add(&acc, &acc, &t);
    }
    *r = acc;
}

/* ---- ark-ec 0.4.2 msm_bigint_wnaf restated ---- */
static size_t ark_window(size_t n) {
    if (n < 32) return 3;
    size_t lg = 63 - (size_t)__builtin_clzll((unsigned long long)n);
    return lg * 69 / 100 + 2;   /* ln_without_floats(n) + 2 */
}
static void make_digits(const uint64_t *scalar, size_t w, size_t num_bits, int64_t *digits, size_t count) {
    uint64_t radix = 1ULL << w, mask = radix - 1, carry = 0;
    (void)num_bits;
    for (size_t i = 0; i < count; i++) {
        size_t bit_offset = i * w, u = bit_offset / 64, b = bit_offset % 64;
        uint64_t buf;
        if (b < 64 - w || u == 3) buf = scalar[u] >> b;
        else buf = (scalar[u] >> b) | (scalar[u + 1] << (64 - b));
        uint64_t coef = carry + (buf & mask);
        carry = (coef + radix / 2) >> w;
        digits[i] = (int64_t)coef - (int64_t)(carry << w);
    }
    digits[count - 1] += (int64_t)(carry << w);
}
static void msm_window(g1j_t *res, const g1a_t *bases, const int64_t *digits, size_t n, size_t count,
                       size_t win, size_t c) {
    size_t nb = (size_t)1 << c;
    g1j_t *buckets = malloc(nb * sizeof *buckets);
    for (size_t i = 0; i < nb; i++) g1j_set_inf(&buckets[i]);
    for (size_t i = 0; i < n; i++) {
        int64_t d = digits[i * count + win];
        if (d > 0) {
            g1j_add_mixed(&buckets[d - 1], &buckets[d - 1], &bases[i]);
        } else if (d < 0) {
            g1a_t nb_ = bases[i];
            fq_neg(&nb_.y, &nb_.y);
            g1j_add_mixed(&buckets[-d - 1], &buckets[-d - 1], &nb_);
        }
    }
    g1j_t running, acc;
    g1j_set_inf(&running);
    g1j_set_inf(&acc);
    for (size_t i = nb; i-- > 0;) {
        g1j_add(&running, &running, &buckets[i]);
        g1j_add(&acc, &acc, &running);
    }
    free(buckets);
    *res = acc;
}
typedef struct {
    g1j_t *wsum; const g1a_t *bases; const int64_t *digits; size_t n, count, c; int unused;
} win_job_t;
static void win_body(void *ctx, size_t w) {
    win_job_t *j = ctx;
    msm_window(&j->wsum[w], j->bases, j->digits, j->n, j->count, w, j->c);
}
void g1_msm_ark_mt(g1j_t *r, const g1a_t *bases, const fr_t *scalars, size_t n, int threads) {
    if (n == 0) { g1j_set_inf(r); return; }
    size_t c = ark_window(n), num_bits = 255, count = (num_bits + c - 1) / c;
    int64_t *digits = malloc(n * count * sizeof *digits);
    for (size_t i = 0; i < n; i++) {
        fr_t k;
        fr_to_canon(&k, &scalars[i]);
        make_digits(k.l, c, num_bits, digits + i * count, count);
    }
    g1j_t *wsum = malloc(count * sizeof *wsum);
    if (threads <= 1) {
        for (size_t w = 0; w < count; w++) msm_window(&wsum[w], bases, digits, n, count, w, c);
    } else {
        /* arkworks' `parallel` feature (NOT enabled by the reference) runs the
         * windows on a rayon pool; this is the generous multi-core CPU arm. */
        win_job_t job = {wsum, bases, digits, n, count, c, 0};
        orc_parallel_for(threads, count, win_body, &job);
    }
    g1j_t total;
    g1j_set_inf(&total);
    for (size_t w = count - 1; w >= 1; w--) {
        g1j_add(&total, &total, &wsum[w]);
        for (size_t k = 0; k < c; k++) g1j_double(&total, &total);
    }
    g1j_add(r, &wsum[0], &total);
    free(wsum);
    free(digits);
}
/* threads used by g1_msm_ark (and so by every protocol function): 1 = what the reference does per party
 * (ark `parallel` is off, dist-primitive/Cargo.toml:18-22); bench.py --impl reference raises it */
static int g_msm_threads = 1;
void orc_set_msm_threads(int t) { g_msm_threads = t < 1 ? 1 : t; }
void g1_msm_ark(g1j_t *r, const g1a_t *bases, const fr_t *scalars, size_t n) {
    g1_msm_ark_mt(r, bases, scalars, n, g_msm_threads);
}

/* ---- vector helpers for the Python side ---- */
typedef struct { const fr_t *k; g1a_t *out; } gen_job_t;
static void gen_body(void *ctx, size_t i) {
    gen_job_t *j = ctx;
    g1j_t g, t;
    g1j_from_affine(&g, &G1_GEN);
    g1j_mul_fr(&t, &g, &j->k[i]);
    g1j_to_affine(&j->out[i], &t);
}
void orc_g1_gen_mul_vec(const fr_t *k, g1a_t *out, size_t n) {
    gen_job_t job = {k, out};
    orc_parallel_for(orc_hw_threads(), n, gen_body, &job);
}
void orc_g1j_to_affine_vec(const g1j_t *in, g1a_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) g1j_to_affine(&out[i], &in[i]);
}
void orc_g1_add_mixed_vec(const g1j_t *a, const g1a_t *b, g1j_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) g1j_add_mixed(&r[i], &a[i], &b[i]);
}
void orc_g1_add_vec(const g1j_t *a, const g1j_t *b, g1j_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) g1j_add(&r[i], &a[i], &b[i]);
}
void orc_g1_double_vec(const g1j_t *a, g1j_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) g1j_double(&r[i], &a[i]);
}
