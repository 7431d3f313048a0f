// Wire formats of the path: what `serialize_compressed` / `deserialize_compressed` put on the reference's TCP links
// (dist-primitive/src/utils/serializing_net.rs:17,25,50,60,88,95).  The NCCL path never serialises -- payloads stay
// in device layout -- but a shim that keeps the reference's own transport needs the bytes, and the byte counters of
// get_comm() are defined in these sizes.
//   Fr  (ark-ff 0.4.2 CanonicalSerialize): 32 B, canonical integer, little endian.
//   G1  (ark-bls12-381 0.4.0 overrides the point encoding with the Zcash / IETF one): 48 B, x big endian; top three
//       bits of byte 0 = compressed (always 1) | infinity | y is the lexicographically larger root.
//       Deserialisation validates: canonical x < p, flags consistent, x^3 + 4 a square, point in the r-torsion.
#include "ctx.h"
#include "g1.cuh"

namespace scz {

__device__ __forceinline__ bool fq_canon_gt(const Fq &a, const Fq &b) {   // canonical integers
    for (int i = 11; i >= 0; i--) {
        if (a.l[i] != b.l[i]) return a.l[i] > b.l[i];
    }
    return false;
}
__device__ __forceinline__ Fq fq_modulus() {
    Fq p;
#pragma unroll
    for (int i = 0; i < 12; i++) p.l[i] = FqP::mod(i);
    return p;
}

__global__ void __launch_bounds__(128) k_g1_serialize(const void *jac, uint8_t *out, size_t n) {
    size_t i = blockIdx.x * (size_t)128 + threadIdx.x;
    if (i >= n) return;
    G1Jac p = g1j_load(jac, i);
    uint8_t *o = out + i * 48;
    if (p.z.is_zero()) {
        o[0] = 0xc0;
        for (int k = 1; k < 48; k++) o[k] = 0;
        return;
    }
    Fq iz = fp_inv(p.z), iz2 = fp_sqr(iz);
    Fq x = fp_to_canon(fp_mul(p.x, iz2));
    Fq ym = fp_mul(p.y, fp_mul(iz2, iz));
    Fq y = fp_to_canon(ym), ny = fp_to_canon(fp_neg(ym));
    uint8_t flags = 0x80 | (fq_canon_gt(y, ny) ? 0x20 : 0);
    for (int k = 0; k < 48; k++) {
        int limb = 11 - k / 4, sh = 24 - 8 * (k % 4);
        o[k] = (uint8_t)(x.l[limb] >> sh);
    }
    o[0] |= flags;
}

// status: 0 ok, 1 malformed (flags / x >= p / not on the curve), 2 on the curve but not in the prime-order subgroup
__global__ void __launch_bounds__(128) k_g1_deserialize(const uint8_t *in, void *jac, uint8_t *status, size_t n) {
    size_t i = blockIdx.x * (size_t)128 + threadIdx.x;
    if (i >= n) return;
    const uint8_t *b = in + i * 48;
    G1Jac out;
    out.x = Fq::one();
    out.y = Fq::one();
    out.z = Fq::zero();
    uint8_t st = 0;
    uint8_t flags = b[0] & 0xe0;
    Fq x;
    for (int limb = 0; limb < 12; limb++) {
        uint32_t v = 0;
        for (int k = 0; k < 4; k++) {
            uint8_t byte = b[(11 - limb) * 4 + k];
            if (limb == 11 && k == 0) byte &= 0x1f;
            v = (v << 8) | byte;
        }
        x.l[limb] = v;
    }
    if (!(flags & 0x80)) st = 1;                       // uncompressed form is not what serialize_compressed writes
    else if (flags & 0x40) {
        if ((flags & 0x20) || !x.is_zero()) st = 1;    // infinity must be 0xc0 00 .. 00
    } else if (!fq_canon_gt(fq_modulus(), x)) st = 1;  // x >= p
    else {
        Fq xm = fp_from_canon(x);
        Fq four = Fq::zero();
        four.l[0] = 4;
        Fq rhs = fp_add(fp_mul(fp_sqr(xm), xm), fp_from_canon(four));
        constexpr uint32_t E[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                                    0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};   // (p + 1) / 4
        Fq y = fp_pow(rhs, E);
        if (fp_sqr(y) != rhs) st = 1;
        else {
            Fq yc = fp_to_canon(y), nyc = fp_to_canon(fp_neg(y));
            bool larger = fq_canon_gt(yc, nyc);
            if (larger != ((flags & 0x20) != 0)) y = fp_neg(y);
            G1X pt;
            pt.x = xm, pt.y = y, pt.zz = Fq::one(), pt.zzz = Fq::one();
            constexpr uint32_t RORD[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                          0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
            if (!g1x_mul_bits(pt, RORD).is_inf()) st = 2;
            else {
                out.x = xm;
                out.y = y;
                out.z = Fq::one();
            }
        }
    }
    status[i] = st;
    g1j_store(jac, i, out);
}

// Fr: canonical little-endian bytes; deserialisation rejects values >= r (status 1)
__global__ void __launch_bounds__(256) k_fr_deserialize(const void *in, void *out, uint8_t *status, size_t n) {
    size_t i = blockIdx.x * (size_t)256 + threadIdx.x;
    if (i >= n) return;
    Fr v = fp_load<FrP>(in, i);
    bool ok = false;
    for (int k = 7; k >= 0; k--) {
        if (v.l[k] != FrP::mod(k)) {
            ok = v.l[k] < FrP::mod(k);
            break;
        }
    }
    status[i] = ok ? 0 : 1;
    fp_store<FrP>(out, i, ok ? fp_from_canon(v) : Fr::zero());
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_g1_serialize_compressed_dev(scz_ctx *h, const void *d_jac, void *d_bytes, size_t n) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (n && (!d_jac || !d_bytes)) return h->c.fail(SCZ_ERR_BAD_ARG, "g1_serialize: null argument");
    if (!n) return SCZ_OK;
    k_g1_serialize<<<ceil_div_u32(n, 128), 128, 0, h->c.stream>>>(d_jac, (uint8_t *)d_bytes, n);
    SCZ_LAUNCH_CHECK(&h->c);
    return SCZ_OK;
}
int32_t scz_g1_deserialize_compressed_dev(scz_ctx *h, const void *d_bytes, void *d_jac, uint8_t *d_status, size_t n) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (n && (!d_jac || !d_bytes || !d_status)) return h->c.fail(SCZ_ERR_BAD_ARG, "g1_deserialize: null argument");
    if (!n) return SCZ_OK;
    k_g1_deserialize<<<ceil_div_u32(n, 128), 128, 0, h->c.stream>>>((const uint8_t *)d_bytes, d_jac, d_status, n);
    SCZ_LAUNCH_CHECK(&h->c);
    return SCZ_OK;
}
int32_t scz_fr_deserialize_dev(scz_ctx *h, const void *d_bytes, void *d_out, uint8_t *d_status, size_t n) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (n && (!d_out || !d_bytes || !d_status)) return h->c.fail(SCZ_ERR_BAD_ARG, "fr_deserialize: null argument");
    if (!n) return SCZ_OK;
    k_fr_deserialize<<<ceil_div_u32(n, 256), 256, 0, h->c.stream>>>(d_bytes, d_out, d_status, n);
    SCZ_LAUNCH_CHECK(&h->c);
    return SCZ_OK;
}

}   // extern "C"
