// Batched Pippenger MSM over BLS12-381 G1 (internal interface).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "ctx.h"

namespace scz {

// one MSM of a batch ("segment"); lives in device memory, read by every kernel of the pipeline
struct MsmSeg {
    const void *bases;     // packed affine, 96 B per point
    const void *scalars;   // Fr Montgomery, 32 B per scalar
    uint32_t len;
    uint32_t point_base;   // prefix sum of len over the batch
    uint32_t c;            // window bits
    uint32_t W;            // windows = ceil(256 / c)
    uint32_t nb;           // buckets per window = 2^(c-1)
    uint32_t bucket_base;  // first global bucket of this segment (window w starts at bucket_base + w*nb)
    uint32_t window_base;  // first global window of this segment
    uint32_t logL;         // bucket-reduction chunk length L = 2^logL = min(nb, 8)
    uint32_t M;            // chunks per window = nb / L (a power of two)
    uint32_t PB;           // log2(M): bit planes of the chunk index
    uint32_t chunk_base;   // first global chunk (window w starts at chunk_base + w*M)
    uint32_t plane_base;   // first global plane slot (window w owns PB + 2 slots from plane_base + w*(PB+2))
    void *out;             // where this segment's result goes (one Jacobian point); null: slot `k` of d_out_jac
};

// d_out: batch Jacobian points (144 B each).  Asynchronous on ctx->stream.
// d_outs (optional, host array of `batch` device pointers) sends each result to its own address instead.
int32_t msm_g1_batched(Ctx *ctx, const void *const *d_bases, const void *const *d_scalars, const size_t *lens,
                       size_t batch, void *d_out_jac, void *const *d_outs = nullptr);
uint32_t msm_pick_window(size_t len);

}   // namespace scz
