/*
 * ORACLE (test infrastructure): the dist-primitive protocols, restated.
 *
 * Two "network" modes, both taken from the reference:
 *   ORC_PARTIES    -- feature `comm` (dist-primitive/src/utils/serializing_net.rs:8-142):
 *                     N = pp.n real parties; gather -> leader closure -> scatter.
 *                     All parties are simulated in this one process: per-party
 *                     inputs are indexed [party], outputs likewise.
 *   ORC_LEADER_SIM -- no `comm` (serializing_net.rs:144-264): a single party
 *                     (the leader); whenever it would receive N messages it sees
 *                     N clones of ITS OWN message (:159-162) and on scatter keeps
 *                     element 0 (:211).  Inputs/outputs are indexed [0] only.
 */
#include <stdlib.h>
#include "oracle.h"

static size_t log2sz(size_t v) {
    size_t l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}
static void msm_any(g1j_t *r, const g1a_t *b, const fr_t *s, size_t n, int use_ark) {
    if (use_ark) g1_msm_ark(r, b, s, n);
    else g1_msm_naive(r, b, s, n);
}

/* ------------------------------------------------------------------ d_msm
 * dist-primitive/src/dmsm.rs:9-43 */
void orc_d_msm(const orc_pp_t *pp, int mode, size_t batch, const size_t *lens, const g1a_t *const *bases,
               const fr_t *const *scalars, g1j_t *out, int use_ark) {
    size_t N = pp->n, P = mode == ORC_PARTIES ? N : 1;
    g1j_t *c = malloc(P * batch * sizeof *c);
    for (size_t p = 0; p < P; p++)
        for (size_t k = 0; k < batch; k++)
            msm_any(&c[p * batch + k], bases[p * batch + k], scalars[p * batch + k], lens[k], use_ark); /* :19-24 */
    g1j_t *sh = malloc(N * sizeof *sh), *sec = malloc(pp->l * sizeof *sec), *pk = malloc(N * sizeof *pk);
    for (size_t k = 0; k < batch; k++) {
        for (size_t j = 0; j < N; j++) sh[j] = c[(mode == ORC_PARTIES ? j : 0) * batch + k];  /* transpose :30 */
        orc_unpack2(pp, 1, sh, sec);                                                        /* :33 */
        g1j_t sum;
        g1j_set_inf(&sum);
        for (size_t i = 0; i < pp->l; i++) g1j_add(&sum, &sum, &sec[i]);                    /* :34 */
        for (size_t i = 0; i < pp->l; i++) sec[i] = sum;                                    /* :35 */
        orc_pack_from_public(pp, 1, sec, pp->l, pk);                                        /* :36 */
        for (size_t p = 0; p < P; p++) out[p * batch + k] = pk[p];                          /* scatter */
    }
    free(c);
    free(sh);
    free(sec);
    free(pk);
}

/* ------------------------------------------------------------------ sumcheck
 * One round of dist-primitive/src/dsumcheck.rs:37-85 (identical loop bodies at
 * :167-219, :227-279, :377-429, :452-504).  Folds f,g in place to len/2. */
static void product_round(fr_t *f, fr_t *g, size_t len, const fr_t *ch, fr3_t *res) {
    size_t h = len / 2;
    fr_t omc, two, s0, s1, s2, t, u, v;
    fr_sub(&omc, &fr_R1, ch);
    fr_from_u64(&two, 2);
    fr_set_zero(&s0);
    fr_set_zero(&s1);
    fr_set_zero(&s2);
    for (size_t i = 0; i < h; i++) {
        fr_mul(&t, &f[i], &g[i]);
        fr_add(&s0, &s0, &t);
        fr_mul(&t, &f[h + i], &g[h + i]);
        fr_add(&s1, &s1, &t);
        fr_mul(&u, &f[h + i], &two);
        fr_sub(&u, &u, &f[i]);            /* -x + y*2 */
        fr_mul(&v, &g[h + i], &two);
        fr_sub(&v, &v, &g[i]);
        fr_mul(&t, &u, &v);
        fr_add(&s2, &s2, &t);
    }
    res->a = s0;
    res->b = s1;
    res->c = s2;
    for (size_t i = 0; i < h; i++) {
        fr_mul(&t, &f[i], &omc);
        fr_mul(&u, &f[h + i], ch);
        fr_add(&f[i], &t, &u);            /* a*(1-r) + b*r */
        fr_mul(&t, &g[i], &omc);
        fr_mul(&u, &g[h + i], ch);
        fr_add(&g[i], &t, &u);
    }
}
static fr_t *dupv(const fr_t *v, size_t n) {
    fr_t *r = malloc((n ? n : 1) * sizeof *r);
    memcpy(r, v, n * sizeof *r);
    return r;
}
/* dsumcheck.rs:28-90 */
size_t orc_sumcheck_product(const fr_t *f, const fr_t *g, size_t len, const fr_t *challenge, fr3_t *out) {
    size_t n = log2sz(len);
    fr_t *ff = dupv(f, len), *gg = dupv(g, len);
    for (size_t i = 0; i < n; i++) product_round(ff, gg, len >> i, &challenge[i], &out[i]);
    fr_set_zero(&out[n].a);
    fr_mul(&out[n].b, &ff[0], &gg[0]);
    fr_set_zero(&out[n].c);
    free(ff);
    free(gg);
    return n + 1;
}

/* ---- single-MLE sumchecks: dsumcheck.rs:6-26, 92-146, 287-357 ---- */
static void sum_round(fr_t *f, size_t len, const fr_t *ch, fr2_t *res) {
    size_t h = len / 2;
    fr_t one, omc, t, u, s0, s1;
    fr_set_one(&one);
    fr_sub(&omc, &one, ch);                   /* F::ONE - challenge[i] */
    fr_set_zero(&s0);
    fr_set_zero(&s1);
    for (size_t i = 0; i < h; i++) {
        fr_add(&s0, &s0, &f[i]);
        fr_add(&s1, &s1, &f[h + i]);
    }
    res->a = s0;
    res->b = s1;
    for (size_t i = 0; i < h; i++) {
        fr_mul(&t, &f[i], &omc);
        fr_mul(&u, &f[h + i], ch);
        fr_add(&f[i], &t, &u);                /* a*(1-r) + b*r */
    }
}
/* dsumcheck.rs:6-26 */
size_t orc_sumcheck(const fr_t *f, size_t len, const fr_t *challenge, fr2_t *out) {
    size_t n = log2sz(len);
    fr_t *ff = dupv(f, len);
    for (size_t i = 0; i < n; i++) sum_round(ff, len >> i, &challenge[i], &out[i]);
    fr_set_zero(&out[n].a);
    out[n].b = ff[0];
    free(ff);
    return n + 1;
}
/* dsumcheck.rs:92-146.  out is [P][n + log2(l) + 1]. */
size_t orc_c_sumcheck(const orc_pp_t *pp, int mode, const fr_t *const *f, size_t len, const fr_t *challenge, fr2_t *out) {
    size_t N = pp->n, P = mode == ORC_PARTIES ? N : 1, l = pp->l;
    size_t n = log2sz(len), ll = log2sz(l), cnt = n + ll + 1;
    fr_t *last = malloc(P * sizeof *last);
    for (size_t p = 0; p < P; p++) {                                   /* Phase 1 :105-121 */
        fr_t *ff = dupv(f[p], len);
        for (size_t i = 0; i < n; i++) sum_round(ff, len >> i, &challenge[i], &out[p * cnt + i]);
        last[p] = ff[0];
        free(ff);
    }
    fr_t *f2 = malloc(P * l * sizeof *f2);
    orc_pss2ss(pp, mode, last, f2);                                    /* :124 */
    for (size_t p = 0; p < P; p++) {
        fr_t *ff = f2 + p * l;
        for (size_t i = 0; i < ll; i++) sum_round(ff, l >> i, &challenge[i], &out[p * cnt + n + i]);   /* :127-141 */
        fr_set_zero(&out[p * cnt + n + ll].a);
        out[p * cnt + n + ll].b = ff[0];                               /* :143 */
    }
    free(last);
    free(f2);
    return cnt;
}
/* dsumcheck.rs:287-357; only the leader's result (workers return an empty Vec, :351-353) */
size_t orc_d_sumcheck(int mode, size_t nparties, const fr_t *const *f, size_t len, const fr_t *challenge, fr2_t *out) {
    size_t N = nparties, P = mode == ORC_PARTIES ? N : 1;
    size_t n = log2sz(len), s = log2sz(N);
    fr2_t *local = malloc(P * (n + 1) * sizeof *local);
    for (size_t p = 0; p < P; p++) {
        fr_t *ff = dupv(f[p], len);
        for (size_t i = 0; i < n; i++) sum_round(ff, len >> i, &challenge[i], &local[p * (n + 1) + i]);
        fr_set_zero(&local[p * (n + 1) + n].a);
        local[p * (n + 1) + n].b = ff[0];                              /* :318 */
        free(ff);
    }
    fr_t *lf = malloc(N * sizeof *lf);
    for (size_t i = 0; i < n; i++) {                                   /* :323-331 */
        fr_set_zero(&out[i].a);
        fr_set_zero(&out[i].b);
        for (size_t j = 0; j < N; j++) {
            const fr2_t *x = &local[(mode == ORC_PARTIES ? j : 0) * (n + 1) + i];
            fr_add(&out[i].a, &out[i].a, &x->a);
            fr_add(&out[i].b, &out[i].b, &x->b);
        }
    }
    for (size_t j = 0; j < N; j++) lf[j] = local[(mode == ORC_PARTIES ? j : 0) * (n + 1) + n].b;   /* :332 */
    for (size_t i = 0; i < s; i++) sum_round(lf, N >> i, &challenge[n + i], &out[n + i]);           /* :334-347 */
    free(local);
    free(lf);
    return n + s;
}

/* unpack.rs:72-97 */
void orc_pss2ss(const orc_pp_t *pp, int mode, const fr_t *share_per_party, fr_t *out /* [P][l] */) {
    size_t N = pp->n, P = mode == ORC_PARTIES ? N : 1, l = pp->l;
    fr_t *sh = malloc(N * sizeof *sh), *sec = malloc(l * sizeof *sec), *pk = malloc(N * sizeof *pk);
    for (size_t j = 0; j < N; j++) sh[j] = share_per_party[mode == ORC_PARTIES ? j : 0];
    orc_unpack(pp, 0, sh, sec);
    for (size_t i = 0; i < l; i++) {
        orc_pack_single(pp, 0, &sec[i], pk);
        for (size_t p = 0; p < P; p++) out[p * l + i] = pk[p];    /* transpose + scatter */
    }
    free(sh);
    free(sec);
    free(pk);
}
/* degree_reduce.rs:29-41 */
void orc_degree_reduce(const orc_pp_t *pp, int mode, const fr_t *share_per_party, fr_t *out /* [P] */) {
    size_t N = pp->n, P = mode == ORC_PARTIES ? N : 1, l = pp->l;
    fr_t *sh = malloc(N * sizeof *sh), *sec = malloc(l * sizeof *sec), *pk = malloc(N * sizeof *pk);
    for (size_t j = 0; j < N; j++) sh[j] = share_per_party[mode == ORC_PARTIES ? j : 0];
    orc_unpack2(pp, 0, sh, sec);
    orc_pack_from_public(pp, 0, sec, l, pk);
    for (size_t p = 0; p < P; p++) out[p] = pk[p];
    free(sh);
    free(sec);
    free(pk);
}

/* dsumcheck.rs:148-285.  out is [P][n + log2(l) + 1]. */
size_t orc_c_sumcheck_product(const orc_pp_t *pp, int mode, const fr_t *const *f, const fr_t *const *g,
                              size_t len, const fr_t *challenge, fr3_t *out) {
    size_t N = pp->n, P = mode == ORC_PARTIES ? N : 1, l = pp->l;
    size_t n = log2sz(len), ll = log2sz(l), cnt = n + ll + 1;
    fr_t *lastf = malloc(P * sizeof *lastf), *lastg = malloc(P * sizeof *lastg);
    for (size_t p = 0; p < P; p++) {                                   /* Phase 1 :167-219 */
        fr_t *ff = dupv(f[p], len), *gg = dupv(g[p], len);
        for (size_t i = 0; i < n; i++) product_round(ff, gg, len >> i, &challenge[i], &out[p * cnt + i]);
        lastf[p] = ff[0];
        lastg[p] = gg[0];
        free(ff);
        free(gg);
    }
    fr_t *f2 = malloc(P * l * sizeof *f2), *g2 = malloc(P * l * sizeof *g2);
    orc_pss2ss(pp, mode, lastf, f2);                                   /* :224 */
    orc_pss2ss(pp, mode, lastg, g2);                                   /* :225 */
    for (size_t p = 0; p < P; p++) {
        fr_t *ff = f2 + p * l, *gg = g2 + p * l;
        for (size_t i = 0; i < ll; i++)                                /* Phase 2 uses challenge[i] (:230) */
            product_round(ff, gg, l >> i, &challenge[i], &out[p * cnt + n + i]);
        fr3_t *fin = &out[p * cnt + n + ll];
        fr_set_zero(&fin->a);
        fr_mul(&fin->b, &ff[0], &gg[0]);                               /* :282 */
        fr_set_zero(&fin->c);
    }
    free(lastf);
    free(lastg);
    free(f2);
    free(g2);
    return cnt;
}

/* dsumcheck.rs:359-512; only the leader's result is produced (workers return an empty Vec, :507-509). */
size_t orc_d_sumcheck_product(int mode, size_t nparties, const fr_t *const *f, const fr_t *const *g, size_t len,
                              const fr_t *challenge, fr3_t *out) {
    size_t N = nparties, P = mode == ORC_PARTIES ? N : 1;
    size_t n = log2sz(len), s = log2sz(N);
    fr3_t *local = malloc(P * (n + 1) * sizeof *local);
    for (size_t p = 0; p < P; p++) {
        fr_t *ff = dupv(f[p], len), *gg = dupv(g[p], len);
        for (size_t i = 0; i < n; i++) product_round(ff, gg, len >> i, &challenge[i], &local[p * (n + 1) + i]);
        fr3_t *last = &local[p * (n + 1) + n];
        last->a = gg[0];                                               /* (g_last, f_last, 0) :433 */
        last->b = ff[0];
        fr_set_zero(&last->c);
        free(ff);
        free(gg);
    }
    for (size_t i = 0; i < n; i++) {                                   /* :440-447 */
        fr3_t acc = local[i];
        for (size_t j = 1; j < N; j++) {
            const fr3_t *x = &local[(mode == ORC_PARTIES ? j : 0) * (n + 1) + i];
            fr_add(&acc.a, &acc.a, &x->a);
            fr_add(&acc.b, &acc.b, &x->b);
            fr_add(&acc.c, &acc.c, &x->c);
        }
        out[i] = acc;
    }
    fr_t *lf = malloc(N * sizeof *lf), *lg = malloc(N * sizeof *lg);
    for (size_t j = 0; j < N; j++) {
        const fr3_t *x = &local[(mode == ORC_PARTIES ? j : 0) * (n + 1) + n];
        lf[j] = x->b;                                                  /* :448 */
        lg[j] = x->a;                                                  /* :449 */
    }
    for (size_t i = 0; i < s; i++) product_round(lf, lg, N >> i, &challenge[n + i], &out[n + i]);   /* :452-504 */
    free(local);
    free(lf);
    free(lg);
    return n + s;
}

/* mle.rs:88-104 */
void orc_fix_variable(const fr_t *evals, size_t len, const fr_t *points, size_t npoints, fr_t *out) {
    size_t n = log2sz(len), k = npoints < n ? npoints : n;
    fr_t *v = dupv(evals, len);
    for (size_t i = 0; i < k; i++) {
        size_t h = (len >> i) / 2;
        fr_t omc, t, u;
        fr_sub(&omc, &fr_R1, &points[i]);
        for (size_t j = 0; j < h; j++) {
            fr_mul(&t, &v[j], &omc);
            fr_mul(&u, &v[h + j], &points[i]);
            fr_add(&v[j], &t, &u);
        }
    }
    memcpy(out, v, (len >> k) * sizeof *out);
    free(v);
}

/* ------------------------------------------------------------------ product tree
 * dacc_product.rs:18-23 */
void orc_sub_index(size_t i, size_t *x0, size_t *x1) {
    size_t first_one = 63 - (size_t)__builtin_clzll((unsigned long long)i);
    size_t x = (i & ~((size_t)1 << first_one)) << 1;
    *x0 = x;
    *x1 = x + 1;
}
/* dacc_product.rs:30-39 / :374-381 */
void orc_acc_product_tree(const fr_t *x, size_t len, fr_t *tree) {
    memcpy(tree, x, len * sizeof *tree);
    memcpy(tree + len, x, len * sizeof *tree);
    for (size_t i = len; i < 2 * len - 1; i++) {
        size_t a, b;
        orc_sub_index(i, &a, &b);
        fr_mul(&tree[i], &tree[a], &tree[b]);
    }
    fr_set_zero(&tree[2 * len - 1]);
}
/* dacc_product.rs:365-414 */
void orc_d_acc_product(int mode, size_t nparties, const fr_t *const *inputs, size_t len, fr_t *const *subtrees,
                       fr_t *leader_tree) {
    size_t N = nparties, P = mode == ORC_PARTIES ? N : 1;
    for (size_t p = 0; p < P; p++) orc_acc_product_tree(inputs[p], len, subtrees[p]);
    for (size_t j = 0; j < N; j++)
        leader_tree[j] = subtrees[mode == ORC_PARTIES ? j : 0][2 * len - 1];   /* the forced zero, :381,390 */
    for (size_t i = N; i < 2 * N - 1; i++) {
        size_t a, b;
        orc_sub_index(i, &a, &b);
        fr_mul(&leader_tree[i], &leader_tree[a], &leader_tree[b]);
    }
    fr_set_zero(&leader_tree[2 * N - 1]);
}

/* ------------------------------------------------------------------ PST / multilinear KZG
 * dpoly_comm.rs:37-67 + mature() :141-150 */
void orc_srs_new(orc_srs_t *srs, const g1j_t *g, const fr_t *s, size_t n) {
    srs->levels = n + 1;
    srs->powers_of_g = malloc((n + 1) * sizeof *srs->powers_of_g);
    srs->level_len = malloc((n + 1) * sizeof *srs->level_len);
    g1j_t *cur = malloc(sizeof *cur);
    cur[0] = *g;
    size_t len = 1;
    for (size_t i = 0;; i++) {
        srs->level_len[i] = len;
        srs->powers_of_g[i] = malloc(len * sizeof(g1a_t));
        for (size_t j = 0; j < len; j++) g1j_to_affine(&srs->powers_of_g[i][j], &cur[j]);
        if (i == n) break;
        g1j_t *nx = malloc(2 * len * sizeof *nx);
        fr_t one_minus;
        fr_sub(&one_minus, &fr_R1, &s[n - i - 1]);
        for (size_t j = 0; j < len; j++) {
            g1j_mul_fr(&nx[j], &cur[j], &one_minus);          /* e * (1 - s[n-i-1]) */
            g1j_mul_fr(&nx[len + j], &cur[j], &s[n - i - 1]); /* chained: e * s[n-i-1] */
        }
        free(cur);
        cur = nx;
        len *= 2;
    }
    free(cur);
}
void orc_srs_from_levels(orc_srs_t *srs, size_t levels, g1a_t **levels_ptr, const size_t *level_len) {
    srs->levels = levels;
    srs->powers_of_g = malloc(levels * sizeof *srs->powers_of_g);
    srs->level_len = malloc(levels * sizeof *srs->level_len);
    for (size_t i = 0; i < levels; i++) {
        srs->level_len[i] = level_len[i];
        srs->powers_of_g[i] = malloc(level_len[i] * sizeof(g1a_t));
        memcpy(srs->powers_of_g[i], levels_ptr[i], level_len[i] * sizeof(g1a_t));
    }
}
void orc_srs_free(orc_srs_t *srs) {
    for (size_t i = 0; i < srs->levels; i++) free(srs->powers_of_g[i]);
    free(srs->powers_of_g);
    free(srs->level_len);
}
/* dpoly_comm.rs:237-243 (= d_local_commit :269-275) */
void orc_commit(const orc_srs_t *srs, const fr_t *peval, size_t len, g1j_t *out, int use_ark) {
    size_t level = log2sz(len);
    if (level >= srs->levels || ((size_t)1 << level) != len) abort();   /* the reference asserts */
    msm_any(out, srs->powers_of_g[level], peval, len, use_ark);
}
/* one fold round of dpoly_comm.rs:309-323: q = hi - lo ; r = (1-u)*lo + u*hi (in place) */
static void open_round(fr_t *cur, size_t len, const fr_t *u, fr_t *q) {
    size_t h = len / 2;
    fr_t omu, a, b;
    fr_sub(&omu, &fr_R1, u);
    for (size_t j = 0; j < h; j++) {
        fr_sub(&q[j], &cur[h + j], &cur[j]);
        fr_mul(&a, &omu, &cur[j]);
        fr_mul(&b, u, &cur[h + j]);
        fr_add(&cur[j], &a, &b);
    }
}
/* dpoly_comm.rs:299-325 (= d_local_open :327-353) */
void orc_open(const orc_srs_t *srs, const fr_t *peval, size_t len, const fr_t *point, fr_t *value, g1j_t *proofs,
              int use_ark) {
    size_t n = log2sz(len);
    fr_t *cur = dupv(peval, len), *q = malloc((len / 2 + 1) * sizeof *q);
    for (size_t i = 0; i < n; i++) {
        open_round(cur, len >> i, &point[i], q);
        orc_commit(srs, q, (len >> i) / 2, &proofs[i], use_ark);
    }
    *value = cur[0];
    free(cur);
    free(q);
}
/* dpoly_comm.rs:244-267 ; out [P][batch] */
void orc_c_commit(const orc_srs_t *const *srs, const orc_pp_t *pp, int mode, size_t batch, const size_t *lens,
                  const fr_t *const *pevals, g1j_t *out, int use_ark) {
    size_t P = mode == ORC_PARTIES ? pp->n : 1;
    const g1a_t **bases = malloc(P * batch * sizeof *bases);
    for (size_t p = 0; p < P; p++)
        for (size_t k = 0; k < batch; k++) {
            size_t level = log2sz(lens[k] * pp->l);
            if (level >= srs[p]->levels || ((size_t)1 << level) != lens[k] * pp->l) abort();
            bases[p * batch + k] = srs[p]->powers_of_g[level];
        }
    orc_d_msm(pp, mode, batch, lens, bases, pevals, out, use_ark);
    free(bases);
}
/* dpoly_comm.rs:401-464 ; value [P], proofs [P][n + log2 l]; returns proofs per party */
size_t orc_c_open(const orc_srs_t *const *srs, const orc_pp_t *pp, int mode, const fr_t *const *peval, size_t len,
                  const fr_t *point, fr_t *value, g1j_t *proofs, int use_ark) {
    size_t P = mode == ORC_PARTIES ? pp->n : 1, l = pp->l;
    size_t n = log2sz(len), ll = log2sz(l), cnt = n + ll;
    fr_t **q = malloc(P * (n ? n : 1) * sizeof *q);
    size_t *lens = malloc((n ? n : 1) * sizeof *lens);
    fr_t *last = malloc(P * sizeof *last);
    for (size_t p = 0; p < P; p++) {                                   /* Phase 1 :418-432 */
        fr_t *cur = dupv(peval[p], len);
        for (size_t i = 0; i < n; i++) {
            size_t h = (len >> i) / 2;
            q[p * n + i] = malloc(h * sizeof(fr_t));
            open_round(cur, len >> i, &point[i], q[p * n + i]);
            lens[i] = h;
        }
        last[p] = cur[0];
        free(cur);
    }
    g1j_t *res = malloc(P * (n ? n : 1) * sizeof *res);
    if (n) orc_c_commit(srs, pp, mode, n, lens, (const fr_t *const *)q, res, use_ark);   /* :436 */
    fr_t *r2 = malloc(P * l * sizeof *r2);
    orc_pss2ss(pp, mode, last, r2);                                    /* :439 */
    for (size_t p = 0; p < P; p++) {
        for (size_t i = 0; i < n; i++) proofs[p * cnt + i] = res[p * n + i];
        fr_t *cur = r2 + p * l, *qq = malloc(l * sizeof *qq);
        for (size_t i = 0; i < ll; i++) {                              /* Phase 2 :442-459, point[i] */
            size_t h = (l >> i) / 2;
            open_round(cur, l >> i, &point[i], qq);
            size_t level = log2sz(h * l);
            msm_any(&proofs[p * cnt + n + i], srs[p]->powers_of_g[level], qq, h, use_ark);
        }
        value[p] = cur[0];
        free(qq);
    }
    for (size_t i = 0; i < P * n; i++) free(q[i]);
    free(q);
    free(lens);
    free(last);
    free(res);
    free(r2);
    return cnt;
}
/* dpoly_comm.rs:276-297: every party ends with the same sum */
void orc_d_commit(const orc_srs_t *const *srs, int mode, size_t nparties, const fr_t *const *peval, size_t len,
                  g1j_t *out, int use_ark) {
    size_t N = nparties, P = mode == ORC_PARTIES ? N : 1;
    g1j_t *loc = malloc(P * sizeof *loc), sum;
    for (size_t p = 0; p < P; p++) orc_commit(srs[p], peval[p], len, &loc[p], use_ark);
    g1j_set_inf(&sum);
    for (size_t j = 0; j < N; j++) g1j_add(&sum, &sum, &loc[mode == ORC_PARTIES ? j : 0]);
    *out = sum;
    free(loc);
}
/* dpoly_comm.rs:355-398: leader answer = (root_open.0, root proofs ++ column sums) */
size_t orc_d_open(const orc_srs_t *const *srs, int mode, size_t nparties, const fr_t *const *peval, size_t len,
                  const fr_t *point, size_t npoint, fr_t *value, g1j_t *proofs, int use_ark) {
    size_t N = nparties, P = mode == ORC_PARTIES ? N : 1;
    size_t n = log2sz(len), pl = log2sz(N);
    (void)npoint;
    fr_t *z = malloc(P * sizeof *z);
    g1j_t *pi = malloc(P * (n ? n : 1) * sizeof *pi);
    for (size_t p = 0; p < P; p++) orc_open(srs[p], peval[p], len, point + pl, &z[p], pi + p * n, use_ark);
    fr_t *lz = malloc(N * sizeof *lz);
    for (size_t j = 0; j < N; j++) lz[j] = z[mode == ORC_PARTIES ? j : 0];
    orc_open(srs[0], lz, N, point, value, proofs, use_ark);            /* root_open :377 */
    for (size_t i = 0; i < n; i++) {                                   /* :374-376 */
        g1j_t sum;
        g1j_set_inf(&sum);
        for (size_t j = 0; j < N; j++) g1j_add(&sum, &sum, &pi[(mode == ORC_PARTIES ? j : 0) * n + i]);
        proofs[pl + i] = sum;
    }
    free(z);
    free(pi);
    free(lz);
    return pl + n;
}
