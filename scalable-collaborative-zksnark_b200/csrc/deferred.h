// Deferred MSM execution.
//
// The reference runs `G::msm` wherever a protocol function reaches one (dmsm.rs:23, dpoly_comm.rs:242,274,457):
// ~800 calls per HyperPlonk proof, most of them tiny.  On the GPU a Pippenger launch sequence has a fixed
// latency floor (bucket reduction + the 255-doubling Horner chain, ~2 ms) whatever its size, so the protocol
// layer never runs an MSM on the spot.  It QUEUES the job (bases, scalars, length, where the result goes) and
// registers what has to happen once the result exists (the leader round of d_msm / d_commit / d_open) as a
// continuation.  `run()` then alternates: one batched launch sequence for everything queued, then the
// continuations in registration order (which may queue more -- the root openings of d_open), until nothing is
// pending.  Every challenge of the protocol is known up-front (dhyperplonk.rs:103-109), so no MSM result feeds
// another MSM's input and a whole proof needs two sequences.  Results are the same group elements as in the
// reference's call-by-call order; in parties mode every party registers the same continuations in the same
// order, so the collectives still pair up.
#pragma once
#include <functional>
#include <memory>
#include <vector>

#include "ctx.h"
#include "msm.h"

namespace scz {

struct Deferred {
    Ctx *ctx;
    std::vector<const void *> bases, scalars;
    std::vector<size_t> lens;
    std::vector<void *> outs;
    uint64_t points = 0, entries = 0;                  // queued so far (entries = sum of len * windows)
    std::vector<std::function<int32_t()>> after;       // continuations, run after the next flush
    std::vector<std::shared_ptr<DevTmp>> keep;         // temporaries that must stay alive until run() returns

    explicit Deferred(Ctx *c) : ctx(c) {}
    Deferred(const Deferred &) = delete;
    Deferred &operator=(const Deferred &) = delete;

    // a temporary owned by this object
    int32_t tmp(size_t bytes, DevTmp **out) {
        auto t = std::make_shared<DevTmp>(ctx);
        SCZ_TRY(t->alloc(bytes));
        keep.push_back(t);
        *out = t.get();
        return SCZ_OK;
    }
    // queue out = sum_i scalars[i] * bases[i]; `out` is one Jacobian point (144 B) on the device
    int32_t add_msm(const void *b, const void *s, size_t len, void *out);
    void then(std::function<int32_t()> f) { after.push_back(std::move(f)); }
    int32_t flush_msm();
    int32_t run();
};

}   // namespace scz
