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

struct scz_pp;

namespace scz {

struct Deferred {
    Ctx *ctx;
    std::vector<const void *> bases, scalars;
    std::vector<size_t> lens;
    std::vector<void *> outs;
    std::vector<uint32_t> pre;                         // per job: window of the fixed-base table in `bases`, 0 = plain
    uint64_t points = 0, entries = 0;                  // queued so far (entries = sum of len * windows)
    std::vector<std::function<int32_t()>> after;       // continuations, run after the next flush
    // Leader closures queued by the continuations (one multi-job launch each instead of one launch per call):
    //   PssJob    the d_msm closure (dmsm.rs:31-38) on a gathered buffer, party-major [party][k]
    //   ColsumJob out[r][i] = sum_j in[j * stride + off + i * 144] for r < replicate (d_commit / d_open sums)
    struct PssJob {
        const void *in;
        void *out;
        uint32_t batch, cta_base;
    };
    struct ColsumJob {
        const void *in;
        void *out;
        uint64_t stride, off;
        uint32_t parties, cols, replicate, col_base;
    };
    std::vector<PssJob> pss_jobs;
    const scz_pp *pss_pp = nullptr;
    std::vector<ColsumJob> colsum_jobs;
    std::vector<std::function<int32_t()>> after2;      // run after the round's scatters
    // The leader rounds of one stage travel together: every continuation REGISTERS its gather / scatter and run()
    // issues ONE collective per stage on the concatenated payloads (about 110 star rounds of a proof become 4).
    struct XferReq {
        const void *send;
        void *recv;
        size_t bytes, wire;
    };
    std::vector<XferReq> gathers, scatters;
    std::vector<std::function<int32_t()>> after_gather;   // run once the gathered data is in place (leader-side work)
    std::vector<std::shared_ptr<DevTmp>> keep;         // temporaries that must stay alive until run() returns

    bool early_pending = false;                        // a sequence started by flush_early() has not been joined yet
    size_t early_after = 0, early_after2 = 0;          // continuations registered before the last flush_early()

    explicit Deferred(Ctx *c) : ctx(c) {}
    ~Deferred();
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
    // pre_c != 0: `b` is a fixed-base table built for pre_c-bit windows (srs.cu), laid out [window][point]
    int32_t add_msm(const void *b, const void *s, size_t len, void *out, uint32_t pre_c = 0);
    void then(std::function<int32_t()> f) { after.push_back(std::move(f)); }
    void then2(std::function<int32_t()> f) { after2.push_back(std::move(f)); }
    void then_gathered(std::function<int32_t()> f) { after_gather.push_back(std::move(f)); }
    // worker_send_or_leader_receive_element: recv (n_parties * bytes, party-major) is only used on the leader
    void gather(const void *send, void *recv, size_t bytes, size_t wire) { gathers.push_back(XferReq{send, recv, bytes, wire}); }
    // worker_receive_or_leader_send_element: send (n_parties * bytes) is only used on the leader; runs after the closures
    void scatter(const void *send, void *recv, size_t bytes, size_t wire) { scatters.push_back(XferReq{send, recv, bytes, wire}); }
    int32_t do_gathers();    // protocols.cu
    int32_t do_scatters();
    void add_pss(const scz_pp *pp, const void *in, uint32_t batch, void *out) {
        pss_pp = pp;
        pss_jobs.push_back(PssJob{in, out, batch, 0});
    }
    void add_colsum(const void *in, size_t stride, size_t off, uint32_t parties, uint32_t cols, void *out, uint32_t replicate) {
        colsum_jobs.push_back(ColsumJob{in, out, stride, off, parties, cols, replicate, 0});
    }
    int32_t flush_msm();
    int32_t flush_early();      // start the sequence of what is queued now, without waiting for it (SCZ_MSM_STREAM=1 only)
    int32_t join_early();
    int32_t flush_msm_on_side(bool wait);
    int32_t flush_closures();   // protocols.cu
    int32_t run();
};

}   // namespace scz
