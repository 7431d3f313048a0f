// The G1 side of PolynomialCommitment (dist-primitive/src/dpoly_comm.rs:30-34): `powers_of_g`, one packed affine
// array per level, plus optional fixed-base tables (srs.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector>

struct scz_srs {
    std::vector<const void *> level;   // device pointers, packed affine, 96 B per point
    std::vector<size_t> len;
    std::vector<void *> owned;
    // fixed-base tables (scz_srs_precompute): table[i] holds 2^(c w) * P_j for every window w of a table_c[i]-bit
    // signed-digit recoding, packed affine, laid out [w][j]; null when the level has none
    std::vector<void *> table;
    std::vector<uint32_t> table_c;
    int device = 0;
};
