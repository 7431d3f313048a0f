// Signed-window recoding of a 255-bit scalar for the Pippenger MSM.
//
// The reference's `G::msm` (ark-ec 0.4.2 VariableBaseMSM, call sites
// dist-primitive/src/dmsm.rs:23, dpoly_comm.rs:242,274,457) recodes scalars into
// signed c-bit digits so that a window needs 2^(c-1) buckets.  Any recoding with
// sum_w d_w 2^(c w) = k gives the same group element, so the window size here is
// chosen for the GPU (segment size, bucket-reduction cost), not copied.
//
// Windows: W = ceil(256 / c), i.e. the top window always has a spare bit above
// bit 254 (r < 2^255), so its digit never needs recentring and |d| <= 2^(c-1)
// for every window: bucket index = |d| - 1 in [0, 2^(c-1)).
#pragma once
#include <stdint.h>
#include "field.cuh"

namespace scz {

SCZ_HD constexpr uint32_t msm_num_windows(uint32_t c) { return (256 + c - 1) / c; }

// raw c-bit window `w` of a canonical 256-bit little-endian integer (c <= 24)
SCZ_HD uint32_t msm_raw_window(const uint32_t (&k)[8], uint32_t c, uint32_t w) {
    uint32_t bit = w * c;
    uint32_t limb = bit >> 5, sh = bit & 31;
    if (limb >= 8) return 0;
    uint64_t v = k[limb];
    if (limb + 1 < 8) v |= (uint64_t)k[limb + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

// One recoding step.  Returns the signed digit in (-2^(c-1), 2^(c-1)] and updates carry.
SCZ_HD int32_t msm_signed_digit(const uint32_t (&k)[8], uint32_t c, uint32_t w, uint32_t &carry) {
    uint32_t coef = msm_raw_window(k, c, w) + carry;
    uint32_t half = 1u << (c - 1);
    if (coef > half) {
        carry = 1;
        return (int32_t)coef - (int32_t)(1u << c);
    }
    carry = 0;
    return (int32_t)coef;
}

}   // namespace scz
