// Batched affine G1 additions with one shared inversion (Montgomery's trick) -- the building block of the
// batched-affine bucket accumulation planned in DESIGN.md section 7.  NOT used by libscz.so yet: round 1 only pins
// the arithmetic (tests/emu, tests/test_emu_arith.py::test_g1_batch_affine_add) so that the kernels of the next
// round start from a checked group law.
//
// One affine addition costs 6 field products here (1 for the running product, 2 to peel the inverse of its own
// denominator off the shared one, 1 for the slope, 1 squaring, 1 for y3) against 9.5 for the XYZZ mixed addition of
// g1.cuh; the inversion itself (a 381-bit Fermat chain, ~570 products) is shared by everything the caller batches:
//   phase 1   d_i = x2 - x1 (or 2 y1 when the operands are equal, 1 when a case needs no inverse),
//             prefix[i] = d_0 * ... * d_i                          -> returns the total
//   (caller)  inverts the total -- per thread here, per CTA / per grid in the kernels to come
//   phase 2   walks back: 1/d_i = inv * prefix[i-1], inv *= d_i, then lambda, x3, y3
// Points are the packed affine of g1.cuh (x = y = 0: infinity).  Every exceptional case is handled: identity
// operands, P + P (tangent slope), P + (-P) (infinity; includes the 2-torsion-free y = 0 guard).
#pragma once
#include "g1.cuh"

namespace scz {

enum BatchAddCase : uint8_t { BA_ADD = 0, BA_DOUBLE = 1, BA_TAKE_P = 2, BA_TAKE_Q = 3, BA_INF = 4 };

// the denominator of one addition and which formula it needs
SCZ_HD Fq g1a_batch_denominator(const G1Affine &p, const G1Affine &q, BatchAddCase &kind) {
    if (p.is_inf()) {
        kind = BA_TAKE_Q;
        return Fq::one();
    }
    if (q.is_inf()) {
        kind = BA_TAKE_P;
        return Fq::one();
    }
    Fq d = fp_sub(q.x, p.x);
    if (!d.is_zero()) {
        kind = BA_ADD;
        return d;
    }
    if (p.y == q.y && !p.y.is_zero()) {
        kind = BA_DOUBLE;
        return fp_dbl(p.y);
    }
    kind = BA_INF;
    return Fq::one();
}
// phase 1 over n additions: prefix[i] = product of the denominators 0 .. i; returns prefix[n - 1] (one for n = 0)
SCZ_HD Fq g1a_batch_phase1(const G1Affine *p, const G1Affine *q, int n, Fq *prefix) {
    Fq acc = Fq::one();
    for (int i = 0; i < n; i++) {
        BatchAddCase k;
        acc = fp_mul(acc, g1a_batch_denominator(p[i], q[i], k));
        prefix[i] = acc;
    }
    return acc;
}
// phase 2: inv = (phase-1 total)^-1; out[i] = p[i] + q[i].  `out` may alias neither input.
SCZ_HD void g1a_batch_phase2(const G1Affine *p, const G1Affine *q, int n, const Fq *prefix, Fq inv, G1Affine *out) {
    for (int i = n - 1; i >= 0; i--) {
        BatchAddCase k;
        Fq d = g1a_batch_denominator(p[i], q[i], k);
        Fq dinv = i ? fp_mul(inv, prefix[i - 1]) : inv;   // 1 / d_i
        inv = fp_mul(inv, d);                              // 1 / (d_0 ... d_{i-1})
        G1Affine r;
        if (k == BA_TAKE_P) r = p[i];
        else if (k == BA_TAKE_Q) r = q[i];
        else if (k == BA_INF) {
            r.x = Fq::zero();
            r.y = Fq::zero();
        } else {
            Fq num;
            if (k == BA_ADD) num = fp_sub(q[i].y, p[i].y);
            else {   // tangent: 3 x^2 / 2 y
                Fq xx = fp_sqr(p[i].x);
                num = fp_add(fp_dbl(xx), xx);
            }
            Fq lam = fp_mul(num, dinv);
            r.x = fp_sub(fp_sub(fp_sqr(lam), p[i].x), q[i].x);
            r.y = fp_sub(fp_mul(lam, fp_sub(p[i].x, r.x)), p[i].y);
        }
        out[i] = r;
    }
}

}   // namespace scz
