// The collective call-site layer of the reference, on device buffers.
//
// Mirrors MPCSerializeNet (dist-primitive/src/utils/serializing_net.rs):
//   comm build  (:8-142)   -> CallbackNet: the host supplies gather / scatter /
//                             all_gather over device memory (bench.py and the
//                             tests plug torch.distributed + NCCL in here);
//   no-comm build (:144-264) -> LeaderSimNet: a single party that sees N clones
//                             of its own message and keeps element 0 of a scatter.
// Byte counters reproduce MPCNet::get_comm in the reference's serialised sizes.
#pragma once
#include <stdint.h>
#include "ctx.h"

namespace scz {

struct Net {
    uint32_t n_parties = 1, party_id = 0;
    uint64_t upload = 0, download = 0;
    bool is_leader() const { return party_id == 0; }
    // a leader -> everybody message of `wire` serialised bytes per party that carries no data on this path (the
    // (F::zero(), vec![]) workers get from d_open, dpoly_comm.rs:382-391): counted like the reference counts it
    // (serializing_net.rs:210 / mpc-net/src/multi.rs:378-417), nothing moves
    void count_scatter(size_t wire) {
        if (is_leader()) upload += wire * (n_parties - 1);
        else download += wire;
    }
    virtual ~Net() {}
    // d_recv: n_parties * bytes on the leader (ignored elsewhere)
    virtual int32_t gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) = 0;
    // d_send: n_parties * bytes on the leader; d_recv: bytes on everyone
    virtual int32_t scatter(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) = 0;
    virtual int32_t all_gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) = 0;
    // the "dynamic" variants with a movable hub (serializing_net.rs:41-74, 98-126): every party sends `bytes` to `root`
    // (d_recv: n_parties * bytes there) / `root` sends slice j of d_send to party j.  *got tells the caller whether it
    // received real data (the leader simulator only does when it is the root / never on a foreign scatter).
    virtual int32_t gather_to(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire) = 0;
    virtual int32_t scatter_from(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire,
                                 bool *got) = 0;
    virtual int32_t sync(Ctx *ctx) = 0;
    // point-to-point bytes between two parties (MPCNet::send_to / recv_from, mpc-net/src/lib.rs:55-61): only nets that
    // have a peer-to-peer transport implement them
    virtual int32_t send_to(Ctx *ctx, uint32_t, const void *, size_t) { return ctx_fail_p2p(ctx); }
    virtual int32_t recv_from(Ctx *ctx, uint32_t, void *, size_t) { return ctx_fail_p2p(ctx); }
    static int32_t ctx_fail_p2p(Ctx *ctx);
    // true when non-leaders receive real data on scatter (false in the leader simulator: there are none)
    virtual bool real() const = 0;
};

struct LeaderSimNet : Net {
    int32_t gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) override;
    int32_t scatter(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) override;
    int32_t all_gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) override;
    int32_t gather_to(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire) override;
    int32_t scatter_from(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire,
                         bool *got) override;
    int32_t sync(Ctx *ctx) override;
    bool real() const override { return false; }
};

struct CallbackNet : Net {
    scz_net_vtable vt;
    int32_t gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) override;
    int32_t scatter(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) override;
    int32_t all_gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) override;
    int32_t gather_to(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire) override;
    int32_t scatter_from(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire,
                         bool *got) override;
    int32_t sync(Ctx *ctx) override;
    bool real() const override { return true; }
};

}   // namespace scz
