// Packed secret sharing maps on the device (internal interface).
#pragma once
#include "ctx.h"

struct scz_pp {
    size_t t, l, n;       // secret-sharing/src/pss.rs:38-41
    int device;
    // device-resident matrices of Fr (Montgomery), row-major
    void *d_pack;         // n x 2l : shares = PACK * (secrets zero-padded to 2l)
    void *d_pack_single;  // n      : pack_single(s)[j] = PS[j] * s
    void *d_unpack;       // l x n
    void *d_unpack2;      // l x n
    void *d_dmsm;         // u[n] | p[n] : the d_msm leader closure pack([sum_l unpack2(.)] * l) (dmsm.rs:31-38) is out_j = p_j * sum_i u_i in_i
};

namespace scz {

enum PssMap { PSS_PACK = 0, PSS_PACK_SINGLE = 1, PSS_UNPACK = 2, PSS_UNPACK2 = 3, PSS_DMSM = 4 };

// Applies one of the four linear maps to `batch` vectors.
// Element (b, j) of the input sits at element index b*in_bstride + j*in_jstride,
// output (b, o) at b*out_bstride + o*out_ostride, so the gather / scatter layouts
// of the leader closures (party-major) are consumed without a transpose pass.
int32_t pss_apply(Ctx *ctx, const scz_pp *pp, PssMap map, int kind, const void *d_in, size_t len_in, size_t in_bstride,
                  size_t in_jstride, size_t batch, void *d_out, size_t out_bstride, size_t out_ostride);

// the d_msm leader closure on a list of gathered buffers (Deferred::PssJob array: in, out, batch, -) in one launch
int32_t pss_dmsm_multi(Ctx *ctx, const scz_pp *pp, const void *jobs_host, size_t njobs);

}   // namespace scz
