// Native data plane: the reference's star collectives (mpc-net/src/lib.rs:64-286, typed by
// dist-primitive/src/utils/serializing_net.rs:8-142) carried by NCCL over NVLink, issued by libscz itself on the
// ctx streams -- no host callback, no serialisation.  See include/scz.h ("native data plane") for the model:
// one hub per process / GPU, `per_rank` parties per hub, one ctx and one host thread per party.
//
// A collective among the N = world * per_rank parties:
//   1. every local party publishes its buffer in the hub, records a `ready` event on its stream and waits at the
//      hub's host barrier;
//   2. local party 0 makes its stream wait for the `ready` events, concatenates the local payloads (device copies),
//      issues ONE grouped ncclSend / ncclRecv (or ncclAllGather) for the rank and records `done`;
//   3. after a second barrier every party's stream waits for `done` (and copies its slice out of the staging area for
//      scatters / all-gathers, recording `copied`, which everybody waits for before the buffers may be reused).
// With one party per rank (8 GPUs, l = 1) steps 1 and 3 vanish: a gather is one grouped send / recv on the stream.
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "ctx.h"
#include "net.h"

struct scz_nccl_hub {
    int device = 0;
    uint32_t rank = 0, world = 1, per_rank = 1;
    ncclComm_t comm = nullptr;
    // host barrier among the local parties
    std::mutex m;
    std::condition_variable cv;
    uint32_t waiting = 0;
    uint64_t gen = 0;
    bool broken = false;
    // state of the collective in flight (written before a barrier, read after it)
    std::vector<const void *> slot;
    void *root_recv = nullptr;
    const void *root_send = nullptr;
    const void *stage_ptr = nullptr;     // where the local parties find their slices after `done`
    std::vector<cudaEvent_t> ev_ready, ev_copied;
    cudaEvent_t ev_done = nullptr;
    void *stage = nullptr;               // staging area for concatenated payloads (grown on demand, stream-ordered)
    size_t stage_cap = 0;
    int *d_sync = nullptr;
    std::atomic<uint64_t> calls[4];
    std::atomic<int> ctxs{0};
    bool owned_by_ctx = false;

    bool barrier() {
        if (per_rank == 1) return !broken;
        std::unique_lock<std::mutex> lk(m);
        if (broken) return false;
        uint64_t g = gen;
        if (++waiting == per_rank) {
            waiting = 0;
            gen++;
            cv.notify_all();
            return true;
        }
        cv.wait(lk, [&] { return gen != g || broken; });
        return !broken;
    }
    void abort() {
        std::lock_guard<std::mutex> lk(m);
        broken = true;
        cv.notify_all();
    }
};

namespace scz {

#define SCZ_NCCL(ctx, expr)                                                                                   \
    do {                                                                                                      \
        ncclResult_t r__ = (expr);                                                                            \
        if (r__ != ncclSuccess) {                                                                             \
            hub->abort();                                                                                     \
            return (ctx)->fail(SCZ_ERR_NET, "NCCL error %d (%s) at %s", (int)r__, ncclGetErrorString(r__), #expr); \
        }                                                                                                     \
    } while (0)
#define HUB_BARRIER(ctx)                                                                        \
    do {                                                                                        \
        if (!hub->barrier()) return (ctx)->fail(SCZ_ERR_NET, "nccl hub: aborted (a party failed)"); \
    } while (0)

struct NcclNet : Net {
    scz_nccl_hub *hub = nullptr;
    uint32_t p = 0;   // local index

    bool multi() const { return hub->per_rank > 1; }
    int32_t ready(Ctx *ctx) {
        if (multi()) SCZ_CUDA(ctx, cudaEventRecord(hub->ev_ready[p], ctx->stream));
        return SCZ_OK;
    }
    int32_t wait_ready(Ctx *ctx) {
        if (multi())
            for (cudaEvent_t e : hub->ev_ready) SCZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e, 0));
        return SCZ_OK;
    }
    int32_t done(Ctx *ctx) {
        if (multi()) SCZ_CUDA(ctx, cudaEventRecord(hub->ev_done, ctx->stream));
        return SCZ_OK;
    }
    int32_t wait_done(Ctx *ctx) {
        if (multi() && p != 0) SCZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, hub->ev_done, 0));
        return SCZ_OK;
    }
    int32_t copied_then_wait(Ctx *ctx) {
        if (!multi()) return SCZ_OK;
        SCZ_CUDA(ctx, cudaEventRecord(hub->ev_copied[p], ctx->stream));
        HUB_BARRIER(ctx);
        for (cudaEvent_t e : hub->ev_copied) SCZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e, 0));
        return SCZ_OK;
    }
    // local party 0 only: the staging area holds at least `bytes` (stream-ordered reallocation)
    int32_t stage_reserve(Ctx *ctx, size_t bytes) {
        if (bytes <= hub->stage_cap) return SCZ_OK;
        if (hub->stage) SCZ_CUDA(ctx, cudaFreeAsync(hub->stage, ctx->stream));
        hub->stage = nullptr;
        hub->stage_cap = 0;
        size_t want = align_up(bytes * 2, 1 << 16);
        SCZ_CUDA(ctx, cudaMallocAsync(&hub->stage, want, ctx->stream));
        hub->stage_cap = want;
        return SCZ_OK;
    }
    // local party 0: the rank's payload (per_rank * bytes, local-party-major) as one device buffer
    int32_t concat_local(Ctx *ctx, size_t bytes, const void **src) {
        const uint32_t P = hub->per_rank;
        if (P == 1) {
            *src = hub->slot[0];
            return SCZ_OK;
        }
        SCZ_TRY(stage_reserve(ctx, P * bytes));
        for (uint32_t q = 0; q < P; q++)
            SCZ_CUDA(ctx, cudaMemcpyAsync((char *)hub->stage + q * bytes, hub->slot[q], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        *src = hub->stage;
        return SCZ_OK;
    }

    int32_t gather_to(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire) override {
        if (root == party_id) download += wire * (n_parties - 1);
        else upload += wire;
        const uint32_t P = hub->per_rank, W = hub->world, rr = root / P, rp = root % P;
        hub->slot[p] = d_send;
        if (hub->rank == rr && p == rp) hub->root_recv = d_recv;
        SCZ_TRY(ready(ctx));
        HUB_BARRIER(ctx);
        if (p == 0) {
            SCZ_TRY(wait_ready(ctx));
            hub->calls[0]++;
            if (W == 1) {
                for (uint32_t q = 0; q < P; q++)
                    SCZ_CUDA(ctx, cudaMemcpyAsync((char *)hub->root_recv + q * bytes, hub->slot[q], bytes, cudaMemcpyDeviceToDevice,
                                                  ctx->stream));
            } else {
                const void *src = nullptr;
                SCZ_TRY(concat_local(ctx, bytes, &src));
                const size_t chunk = P * bytes;
                if (hub->rank == rr) {
                    char *dst = (char *)hub->root_recv;
                    SCZ_CUDA(ctx, cudaMemcpyAsync(dst + rr * chunk, src, chunk, cudaMemcpyDeviceToDevice, ctx->stream));
                    SCZ_NCCL(ctx, ncclGroupStart());
                    for (uint32_t r = 0; r < W; r++)
                        if (r != rr) SCZ_NCCL(ctx, ncclRecv(dst + r * chunk, chunk, ncclUint8, (int)r, hub->comm, ctx->stream));
                    SCZ_NCCL(ctx, ncclGroupEnd());
                } else {
                    SCZ_NCCL(ctx, ncclSend(src, chunk, ncclUint8, (int)rr, hub->comm, ctx->stream));
                }
            }
            SCZ_TRY(done(ctx));
        }
        HUB_BARRIER(ctx);
        return wait_done(ctx);   // the root's buffer is filled, every send buffer may be reused
    }
    int32_t scatter_from(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire, bool *got) override {
        if (root == party_id) upload += wire * (n_parties - 1);
        else download += wire;
        if (got) *got = true;
        const uint32_t P = hub->per_rank, W = hub->world, rr = root / P, rp = root % P;
        if (hub->rank == rr && p == rp) hub->root_send = d_send;
        if (p == 0) hub->slot[0] = d_recv;
        SCZ_TRY(ready(ctx));
        HUB_BARRIER(ctx);
        if (p == 0) {
            SCZ_TRY(wait_ready(ctx));
            const size_t chunk = P * bytes;
            if (W == 1) {
                hub->stage_ptr = hub->root_send;
            } else {
                hub->calls[1]++;
                if (hub->rank == rr) {
                    const char *src = (const char *)hub->root_send;
                    SCZ_NCCL(ctx, ncclGroupStart());
                    for (uint32_t r = 0; r < W; r++)
                        if (r != rr) SCZ_NCCL(ctx, ncclSend(src + r * chunk, chunk, ncclUint8, (int)r, hub->comm, ctx->stream));
                    SCZ_NCCL(ctx, ncclGroupEnd());
                    hub->stage_ptr = src + rr * chunk;
                } else if (P == 1) {
                    SCZ_NCCL(ctx, ncclRecv(d_recv, bytes, ncclUint8, (int)rr, hub->comm, ctx->stream));
                    hub->stage_ptr = nullptr;   // already in place
                } else {
                    SCZ_TRY(stage_reserve(ctx, chunk));
                    SCZ_NCCL(ctx, ncclRecv(hub->stage, chunk, ncclUint8, (int)rr, hub->comm, ctx->stream));
                    hub->stage_ptr = hub->stage;
                }
            }
            SCZ_TRY(done(ctx));
        }
        HUB_BARRIER(ctx);
        SCZ_TRY(wait_done(ctx));
        if (hub->stage_ptr)
            SCZ_CUDA(ctx, cudaMemcpyAsync(d_recv, (const char *)hub->stage_ptr + p * bytes, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return copied_then_wait(ctx);   // the staging area / the root's send buffer is free again
    }
    int32_t gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) override {
        return gather_to(ctx, 0, d_send, d_recv, bytes, wire);
    }
    int32_t scatter(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) override {
        return scatter_from(ctx, 0, d_send, d_recv, bytes, wire, nullptr);
    }
    int32_t all_gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) override {
        upload += wire * (n_parties - 1);
        download += wire * (n_parties - 1);
        const uint32_t P = hub->per_rank, W = hub->world;
        hub->slot[p] = d_send;
        if (p == 0) hub->root_recv = d_recv;
        SCZ_TRY(ready(ctx));
        HUB_BARRIER(ctx);
        if (p == 0) {
            SCZ_TRY(wait_ready(ctx));
            hub->calls[2]++;
            if (W == 1) {
                for (uint32_t q = 0; q < P; q++)
                    SCZ_CUDA(ctx, cudaMemcpyAsync((char *)d_recv + q * bytes, hub->slot[q], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            } else {
                const void *src = nullptr;
                SCZ_TRY(concat_local(ctx, bytes, &src));
                SCZ_NCCL(ctx, ncclAllGather(src, d_recv, P * bytes, ncclUint8, hub->comm, ctx->stream));
            }
            SCZ_TRY(done(ctx));
        }
        HUB_BARRIER(ctx);
        SCZ_TRY(wait_done(ctx));
        if (p != 0)
            SCZ_CUDA(ctx, cudaMemcpyAsync(d_recv, hub->root_recv, (size_t)n_parties * bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return copied_then_wait(ctx);   // party 0's buffer has been read by everybody
    }
    // MPCNet::sync (mpc-net/src/lib.rs:275-286): nobody's stream passes this point before everybody's reached it
    int32_t sync(Ctx *ctx) override {
        SCZ_TRY(ready(ctx));
        HUB_BARRIER(ctx);
        if (p == 0) {
            SCZ_TRY(wait_ready(ctx));
            if (hub->world > 1) {
                hub->calls[3]++;
                SCZ_NCCL(ctx, ncclAllReduce(hub->d_sync, hub->d_sync, 1, ncclInt32, ncclSum, hub->comm, ctx->stream));
            }
            SCZ_TRY(done(ctx));
        }
        HUB_BARRIER(ctx);
        return wait_done(ctx);
    }
    // MPCNet::send_to / recv_from as ncclSend / ncclRecv on the ctx stream (one party per rank: the peer IS the rank)
    int32_t send_to(Ctx *ctx, uint32_t peer, const void *d_buf, size_t bytes) override {
        if (hub->per_rank != 1 || peer >= n_parties || peer == party_id || hub->world < 2)
            return ctx->fail(SCZ_ERR_BAD_ARG, "net send_to: needs one party per rank and a peer other than self");
        upload += bytes;
        SCZ_NCCL(ctx, ncclSend(d_buf, bytes, ncclUint8, (int)peer, hub->comm, ctx->stream));
        return SCZ_OK;
    }
    int32_t recv_from(Ctx *ctx, uint32_t peer, void *d_buf, size_t bytes) override {
        if (hub->per_rank != 1 || peer >= n_parties || peer == party_id || hub->world < 2)
            return ctx->fail(SCZ_ERR_BAD_ARG, "net recv_from: needs one party per rank and a peer other than self");
        download += bytes;
        SCZ_NCCL(ctx, ncclRecv(d_buf, bytes, ncclUint8, (int)peer, hub->comm, ctx->stream));
        return SCZ_OK;
    }
    bool real() const override { return true; }
    ~NcclNet() override {
        if (hub) {
            hub->ctxs--;
            if (hub->owned_by_ctx) scz_nccl_hub_destroy(hub);
        }
    }
};

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_nccl_unique_id(void *uid128) {
    if (!uid128) return SCZ_ERR_BAD_ARG;
    static_assert(sizeof(ncclUniqueId) == SCZ_NCCL_UID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return SCZ_ERR_NET;
    memcpy(uid128, &id, sizeof id);
    return SCZ_OK;
}

int32_t scz_nccl_hub_create(int32_t device, uint32_t rank, uint32_t nranks, uint32_t parties_per_rank, const void *uid128,
                            scz_nccl_hub **out) {
    if (!out || !nranks || rank >= nranks || !parties_per_rank || (nranks > 1 && !uid128)) return SCZ_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0, prev = -1;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        fprintf(stderr, "scz_nccl_hub_create: no usable CUDA device %d; libscz has no CPU path\n", device);
        return SCZ_ERR_CUDA;
    }
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return SCZ_ERR_CUDA;
    scz_nccl_hub *h = new scz_nccl_hub();
    h->device = device;
    h->rank = rank, h->world = nranks, h->per_rank = parties_per_rank;
    h->slot.assign(parties_per_rank, nullptr);
    for (auto &c : h->calls) c = 0;
    bool ok = true, nccl_failed = false;
    if (nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, uid128, sizeof id);
        ncclResult_t r = ncclCommInitRank(&h->comm, (int)nranks, id, (int)rank);
        if (r != ncclSuccess) {
            fprintf(stderr, "scz_nccl_hub_create: ncclCommInitRank failed: %s\n", ncclGetErrorString(r));
            ok = false, nccl_failed = true;
        }
    }
    ok = ok && cudaMalloc(&h->d_sync, sizeof(int)) == cudaSuccess && cudaMemset(h->d_sync, 0, sizeof(int)) == cudaSuccess;
    if (ok && parties_per_rank > 1) {
        h->ev_ready.resize(parties_per_rank);
        h->ev_copied.resize(parties_per_rank);
        for (auto &e : h->ev_ready) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
        for (auto &e : h->ev_copied) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming) == cudaSuccess;
    }
    if (prev >= 0) cudaSetDevice(prev);
    if (!ok) {
        scz_nccl_hub_destroy(h);
        return nccl_failed ? SCZ_ERR_NET : SCZ_ERR_CUDA;
    }
    *out = h;
    return SCZ_OK;
}

void scz_nccl_hub_destroy(scz_nccl_hub *h) {
    if (!h) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->comm) ncclCommDestroy(h->comm);
    for (auto e : h->ev_ready) if (e) cudaEventDestroy(e);
    for (auto e : h->ev_copied) if (e) cudaEventDestroy(e);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->stage) cudaFree(h->stage);
    if (h->d_sync) cudaFree(h->d_sync);
    if (prev >= 0) cudaSetDevice(prev);
    delete h;
}

void scz_nccl_hub_abort(scz_nccl_hub *h) {
    if (h) h->abort();
}

int32_t scz_nccl_hub_calls(const scz_nccl_hub *h, uint64_t out4[4]) {
    if (!h || !out4) return SCZ_ERR_BAD_ARG;
    for (int i = 0; i < 4; i++) out4[i] = h->calls[i].load();
    return SCZ_OK;
}

int32_t scz_ctx_create_on_hub(scz_nccl_hub *hub, uint32_t local_index, scz_ctx **out) {
    if (!hub || !out || local_index >= hub->per_rank) return SCZ_ERR_BAD_ARG;
    // a ctx with the leader simulator first (party 0 of 1 is always valid), then swap the net in
    scz_ctx *h = nullptr;
    int32_t rc = scz_ctx_create(hub->device, 0, 1, nullptr, &h);
    if (rc != SCZ_OK) return rc;
    delete h->c.net;
    NcclNet *n = new NcclNet();
    n->hub = hub;
    n->p = local_index;
    n->n_parties = hub->world * hub->per_rank;
    n->party_id = hub->rank * hub->per_rank + local_index;
    hub->ctxs++;
    h->c.net = n;
    *out = h;
    return SCZ_OK;
}

int32_t scz_ctx_create_nccl(int32_t device, uint32_t rank, uint32_t nranks, const void *uid128, scz_ctx **out) {
    if (!out) return SCZ_ERR_BAD_ARG;
    scz_nccl_hub *hub = nullptr;
    int32_t rc = scz_nccl_hub_create(device, rank, nranks, 1, uid128, &hub);
    if (rc != SCZ_OK) return rc;
    rc = scz_ctx_create_on_hub(hub, 0, out);
    if (rc != SCZ_OK) {
        scz_nccl_hub_destroy(hub);
        return rc;
    }
    hub->owned_by_ctx = true;
    return SCZ_OK;
}

int32_t scz_net_gather(scz_ctx *h, uint32_t root, const void *d_send, void *d_recv, size_t bytes) {
    scz::DeviceGuard dg__(h);
    if (!h || !d_send || root >= h->c.net->n_parties) return SCZ_ERR_BAD_ARG;
    return h->c.net->gather_to(&h->c, root, d_send, d_recv, bytes, bytes);
}
int32_t scz_net_scatter(scz_ctx *h, uint32_t root, const void *d_send, void *d_recv, size_t bytes) {
    scz::DeviceGuard dg__(h);
    if (!h || !d_recv || root >= h->c.net->n_parties) return SCZ_ERR_BAD_ARG;
    bool got = false;
    return h->c.net->scatter_from(&h->c, root, d_send, d_recv, bytes, bytes, &got);
}
int32_t scz_net_all_gather(scz_ctx *h, const void *d_send, void *d_recv, size_t bytes) {
    scz::DeviceGuard dg__(h);
    if (!h || !d_send || !d_recv) return SCZ_ERR_BAD_ARG;
    return h->c.net->all_gather(&h->c, d_send, d_recv, bytes, bytes);
}
int32_t scz_net_send(scz_ctx *h, uint32_t peer, const void *d_buf, size_t bytes) {
    scz::DeviceGuard dg__(h);
    if (!h || (bytes && !d_buf)) return SCZ_ERR_BAD_ARG;
    return h->c.net->send_to(&h->c, peer, d_buf, bytes);
}
int32_t scz_net_recv(scz_ctx *h, uint32_t peer, void *d_buf, size_t bytes) {
    scz::DeviceGuard dg__(h);
    if (!h || (bytes && !d_buf)) return SCZ_ERR_BAD_ARG;
    return h->c.net->recv_from(&h->c, peer, d_buf, bytes);
}
int32_t scz_net_sync(scz_ctx *h) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    return h->c.net->sync(&h->c);
}

}   // extern "C"
