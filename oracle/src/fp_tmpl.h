/*
 * ORACLE (test infrastructure, never shipped): prime-field arithmetic template.
 *
 * Restates the arithmetic the reference gets from the un-vendored crate
 * ark-ff 0.4.2 (Cargo.lock:136-137): Fp<MontBackend<_, N>, N> = N little-endian
 * u64 limbs holding a*R mod p, R = 2^(64N).  Instantiated twice by fields.c:
 *   fr_*  N=4  BLS12-381 scalar field   (every `F`/`E::ScalarField` in dist-primitive)
 *   fq_*  N=6  BLS12-381 base field     (coordinates of `E::G1`)
 * PARITY UNPINNED at the byte level: arkworks is not in /root/reference; pinned
 * instead by public curve constants, the Python big-int twin (oracle/py_twin.py)
 * and the reference's own property tests.
 *
 * Include with FP(x) and NL defined.
 */
#include <stdint.h>
#include <string.h>

typedef struct { uint64_t l[NL]; } FP(t);

extern FP(t) FP(MOD);      /* modulus (canonical integer) */
extern FP(t) FP(R1);       /* R mod p  = Montgomery one */
extern FP(t) FP(R2);       /* R^2 mod p */
extern uint64_t FP(INV);   /* -p^{-1} mod 2^64 */

static inline int FP(is_zero)(const FP(t) *a) {
    uint64_t acc = 0;
    for (int i = 0; i < NL; i++) acc |= a->l[i];
    return acc == 0;
}
static inline int FP(eq)(const FP(t) *a, const FP(t) *b) {
    uint64_t acc = 0;
    for (int i = 0; i < NL; i++) acc |= a->l[i] ^ b->l[i];
    return acc == 0;
}
/* a >= b on raw integers */
static inline int FP(geq_raw)(const FP(t) *a, const FP(t) *b) {
    for (int i = NL - 1; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static inline uint64_t FP(add_raw)(FP(t) *r, const FP(t) *a, const FP(t) *b) {
    unsigned __int128 c = 0;
    for (int i = 0; i < NL; i++) {
        c += (unsigned __int128)a->l[i] + b->l[i];
        r->l[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline uint64_t FP(sub_raw)(FP(t) *r, const FP(t) *a, const FP(t) *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < NL; i++) {
        unsigned __int128 d = (unsigned __int128)a->l[i] - b->l[i] - borrow;
        r->l[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
static inline void FP(add)(FP(t) *r, const FP(t) *a, const FP(t) *b) {
    FP(t) s;
    uint64_t c = FP(add_raw)(&s, a, b);
    if (c || FP(geq_raw)(&s, &FP(MOD))) FP(sub_raw)(&s, &s, &FP(MOD));
    *r = s;
}
static inline void FP(sub)(FP(t) *r, const FP(t) *a, const FP(t) *b) {
    FP(t) s;
    if (FP(sub_raw)(&s, a, b)) FP(add_raw)(&s, &s, &FP(MOD));
    *r = s;
}
static inline void FP(neg)(FP(t) *r, const FP(t) *a) {
    if (FP(is_zero)(a)) { *r = *a; return; }
    FP(sub_raw)(r, &FP(MOD), a);
}
static inline void FP(dbl)(FP(t) *r, const FP(t) *a) { FP(add)(r, a, a); }

/* Montgomery product a*b/R mod p: CIOS with the two inner passes fused (one sweep over j carries both the a*b_i
 * row and the m*p row).  Valid because the top bit of both moduli is clear (Fr: 255 of 256 bits, Fq: 381 of 384),
 * so the running sum never needs limb N + 1 -- the "no-carry" form ark-ff 0.4.2's MontBackend uses for these
 * moduli (montgomery_backend.rs, `can_use_no_carry_mul_optimization`) [ark-recall]. */
static inline void FP(mul)(FP(t) *r, const FP(t) *a, const FP(t) *b) {
    uint64_t t[NL];
    memset(t, 0, sizeof t);
    for (int i = 0; i < NL; i++) {
        const uint64_t bi = b->l[i];
        unsigned __int128 c1 = (unsigned __int128)a->l[0] * bi + t[0];
        const uint64_t m = (uint64_t)c1 * FP(INV);
        unsigned __int128 c2 = (unsigned __int128)m * FP(MOD).l[0] + (uint64_t)c1;
        c1 >>= 64;
        c2 >>= 64;
        for (int j = 1; j < NL; j++) {
            c1 += (unsigned __int128)a->l[j] * bi + t[j];
            c2 += (unsigned __int128)m * FP(MOD).l[j] + (uint64_t)c1;
            t[j - 1] = (uint64_t)c2;
            c1 >>= 64;
            c2 >>= 64;
        }
        t[NL - 1] = (uint64_t)c1 + (uint64_t)c2;
    }
    /* branch-free final subtraction: keep t - p unless it borrowed */
    FP(t) s, d;
    memcpy(s.l, t, sizeof s.l);
    uint64_t keep = 0 - FP(sub_raw)(&d, &s, &FP(MOD));   /* all ones when t < p */
    for (int i = 0; i < NL; i++) r->l[i] = (s.l[i] & keep) | (d.l[i] & ~keep);
}
static inline void FP(sqr)(FP(t) *r, const FP(t) *a) { FP(mul)(r, a, a); }

/* Montgomery form <-> canonical integer (ark-ff `into_bigint` / `from_bigint`). */
static inline void FP(to_canon)(FP(t) *r, const FP(t) *a) {
    FP(t) one;
    memset(&one, 0, sizeof one);
    one.l[0] = 1;
    FP(mul)(r, a, &one);
}
static inline void FP(from_canon)(FP(t) *r, const FP(t) *a) { FP(mul)(r, a, &FP(R2)); }
static inline void FP(from_u64)(FP(t) *r, uint64_t v) {
    FP(t) x;
    memset(&x, 0, sizeof x);
    x.l[0] = v;
    FP(from_canon)(r, &x);
}
static inline void FP(set_zero)(FP(t) *r) { memset(r, 0, sizeof *r); }
static inline void FP(set_one)(FP(t) *r) { *r = FP(R1); }

/* a^e, e = canonical little-endian limbs */
static inline void FP(pow)(FP(t) *r, const FP(t) *a, const uint64_t *e, int elimbs) {
    FP(t) acc = FP(R1), base = *a;
    for (int i = 0; i < elimbs * 64; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) FP(mul)(&acc, &acc, &base);
        FP(sqr)(&base, &base);
    }
    *r = acc;
}

/* Field inverse by binary extended Euclid, the algorithm ark-ff 0.4 uses
 * (Guajardo-Kumar-Paar-Pelzl alg. 16): b starts at R^2 so the result is
 * already in Montgomery form.  Returns 0 for a == 0 (arkworks: None -> the
 * reference's `a / b` panics, hyperplonk/src/dhyperplonk.rs:339). */
static inline int FP(inv)(FP(t) *r, const FP(t) *a) {
    if (FP(is_zero)(a)) return 0;
    FP(t) one;
    memset(&one, 0, sizeof one);
    one.l[0] = 1;
    FP(t) u = *a, v = FP(MOD), b = FP(R2), c;
    memset(&c, 0, sizeof c);
    while (!FP(eq)(&u, &one) && !FP(eq)(&v, &one)) {
        while ((u.l[0] & 1) == 0) {
            for (int i = 0; i < NL - 1; i++) u.l[i] = (u.l[i] >> 1) | (u.l[i + 1] << 63);
            u.l[NL - 1] >>= 1;
            uint64_t carry = 0;
            if (b.l[0] & 1) carry = FP(add_raw)(&b, &b, &FP(MOD));
            for (int i = 0; i < NL - 1; i++) b.l[i] = (b.l[i] >> 1) | (b.l[i + 1] << 63);
            b.l[NL - 1] = (b.l[NL - 1] >> 1) | (carry << 63);
        }
        while ((v.l[0] & 1) == 0) {
            for (int i = 0; i < NL - 1; i++) v.l[i] = (v.l[i] >> 1) | (v.l[i + 1] << 63);
            v.l[NL - 1] >>= 1;
            uint64_t carry = 0;
            if (c.l[0] & 1) carry = FP(add_raw)(&c, &c, &FP(MOD));
            for (int i = 0; i < NL - 1; i++) c.l[i] = (c.l[i] >> 1) | (c.l[i + 1] << 63);
            c.l[NL - 1] = (c.l[NL - 1] >> 1) | (carry << 63);
        }
        if (FP(geq_raw)(&u, &v)) {  /* v <= u */
            FP(sub_raw)(&u, &u, &v);
            FP(sub)(&b, &b, &c);
        } else {
            FP(sub_raw)(&v, &v, &u);
            FP(sub)(&c, &c, &b);
        }
    }
    *r = FP(eq)(&u, &one) ? b : c;
    return 1;
}

/* Derive R, R^2, INV from the modulus (called once from orc_init). */
static inline void FP(derive_constants)(void) {
    uint64_t p0 = FP(MOD).l[0], x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p0 * x;   /* Newton: x = p0^{-1} mod 2^64 */
    FP(INV) = (uint64_t)0 - x;
    FP(t) acc;
    memset(&acc, 0, sizeof acc);
    acc.l[0] = 1;
    for (int i = 0; i < 64 * NL; i++) FP(add)(&acc, &acc, &acc);
    FP(R1) = acc;
    for (int i = 0; i < 64 * NL; i++) FP(add)(&acc, &acc, &acc);
    FP(R2) = acc;
}
