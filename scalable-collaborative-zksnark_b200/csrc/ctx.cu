// Context, device memory, and the two network back-ends.
#include <stdarg.h>
#include <atomic>
#include <stdlib.h>
#include <string.h>

#include "ctx.h"
#include "net.h"

namespace scz {

static std::atomic<int> g_live_ctx[64];
int ctx_live_on_device(int device) { return device >= 0 && device < 64 ? g_live_ctx[device].load() : 1; }

int32_t Ctx::fail(int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
}
int32_t Ctx::cuda(cudaError_t e, const char *what) {
    return fail(SCZ_ERR_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}
int32_t Ctx::pinned_reserve(size_t bytes) {
    if (bytes <= pinned_cap) return SCZ_OK;
    if (pinned) SCZ_CUDA(this, cudaFreeHost(pinned));
    pinned = nullptr;
    pinned_cap = 0;
    size_t want = align_up(bytes, 1 << 16);
    SCZ_CUDA(this, cudaMallocHost(&pinned, want));
    pinned_cap = want;
    return SCZ_OK;
}

int32_t Ctx::h2d_staged(void *d_dst, const void *h_src, size_t bytes) {
    if (!bytes) return SCZ_OK;
    if (bytes > STAGE_BYTES) {   // rare (tens of thousands of segments): pageable copy, which the runtime stages itself
        SCZ_CUDA(this, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
        SCZ_CUDA(this, cudaStreamSynchronize(stream));
        return SCZ_OK;
    }
    if (!stage[0].p) {   // all slots at once (see ctx.h)
        char *block = nullptr;
        SCZ_CUDA(this, cudaMallocHost(&block, STAGE_BYTES * STAGE_SLOTS));
        for (int i = 0; i < STAGE_SLOTS; i++) {
            stage[i].p = block + (size_t)i * STAGE_BYTES;
            SCZ_CUDA(this, cudaEventCreateWithFlags(&stage[i].ev, cudaEventDisableTiming));
        }
    }
    Stage &s = stage[stage_next];
    stage_next = (stage_next + 1) % STAGE_SLOTS;
    if (s.busy) SCZ_CUDA(this, cudaEventSynchronize(s.ev));
    memcpy(s.p, h_src, bytes);
    SCZ_CUDA(this, cudaMemcpyAsync(d_dst, s.p, bytes, cudaMemcpyHostToDevice, stream));
    SCZ_CUDA(this, cudaEventRecord(s.ev, stream));
    s.busy = true;
    return SCZ_OK;
}

void Ctx::prof_begin(int id) {
    ProfRec r;
    r.id = id;
    for (cudaEvent_t *e : {&r.a, &r.b}) {
        if (!prof_pool.empty()) {
            *e = prof_pool.back();
            prof_pool.pop_back();
        } else {
            cudaEventCreate(e);
        }
    }
    cudaEventRecord(r.a, stream);
    prof_recs.push_back(r);
}
void Ctx::prof_end() {
    if (!prof_recs.empty()) cudaEventRecord(prof_recs.back().b, stream);
}
void Ctx::prof_clear() {
    for (ProfRec &r : prof_recs) {
        prof_pool.push_back(r.a);
        prof_pool.push_back(r.b);
    }
    prof_recs.clear();
}

// dst[j] = src for j < copies, 16-byte words (every payload of the path is a multiple of 16 B)
__global__ void k_replicate16(uint4 *dst, const uint4 *src, uint32_t words, uint32_t copies) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= words) return;
    uint4 v = src[i];
    for (uint32_t j = 0; j < copies; j++) dst[(size_t)j * words + i] = v;
}
static int32_t replicate(Ctx *ctx, void *d_recv, const void *d_send, size_t bytes, uint32_t copies) {
    if (bytes % 16 == 0 && bytes / 16 < (1u << 31) && ((uintptr_t)d_recv % 16) == 0 && ((uintptr_t)d_send % 16) == 0) {
        uint32_t words = (uint32_t)(bytes / 16);
        if (!words) return SCZ_OK;
        k_replicate16<<<ceil_div_u32(words, 256), 256, 0, ctx->stream>>>((uint4 *)d_recv, (const uint4 *)d_send, words, copies);
        SCZ_LAUNCH_CHECK(ctx);
        return SCZ_OK;
    }
    for (uint32_t j = 0; j < copies; j++)
        SCZ_CUDA(ctx, cudaMemcpyAsync((char *)d_recv + (size_t)j * bytes, d_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return SCZ_OK;
}

int32_t Net::ctx_fail_p2p(Ctx *ctx) { return ctx->fail(SCZ_ERR_NET, "this net has no point-to-point transport"); }

// ------------------------------------------------------------------ leader simulator
// serializing_net.rs:147-167: the leader "receives" n_parties clones of its own message
int32_t LeaderSimNet::gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) {
    download += wire * (n_parties - 1);
    return replicate(ctx, d_recv, d_send, bytes, n_parties);
}
// serializing_net.rs:192-215: counts the N-1 outgoing messages, keeps element 0
int32_t LeaderSimNet::scatter(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) {
    upload += wire * (n_parties - 1);
    SCZ_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return SCZ_OK;
}
// hyperplonk/src/dhyperplonk.rs:271-294 without `comm`: the own vector is used N times (:289-293)
int32_t LeaderSimNet::all_gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) {
    upload += wire * (n_parties - 1);
    return replicate(ctx, d_recv, d_send, bytes, n_parties);
}
// serializing_net.rs:170-190: only the receiver "receives" (N clones of its own message); everybody else just counts
int32_t LeaderSimNet::gather_to(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire) {
    if (root != party_id) {
        upload += wire;
        return SCZ_OK;
    }
    download += wire * (n_parties - 1);
    return replicate(ctx, d_recv, d_send, bytes, n_parties);
}
// serializing_net.rs:217-241: the hub counts its N - 1 messages and keeps element 0; everybody else gets T::default()
int32_t LeaderSimNet::scatter_from(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire,
                                   bool *got) {
    if (root == party_id) {
        upload += wire * (n_parties - 1);
        SCZ_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (got) *got = true;
    } else {
        SCZ_CUDA(ctx, cudaMemsetAsync(d_recv, 0, bytes, ctx->stream));
        if (got) *got = false;
    }
    return SCZ_OK;
}
int32_t LeaderSimNet::sync(Ctx *) { return SCZ_OK; }

// ------------------------------------------------------------------ host-supplied collectives
int32_t CallbackNet::gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) {
    if (is_leader()) download += wire * (n_parties - 1);
    else upload += wire;
    if (vt.gather(vt.user, d_send, d_recv, bytes, wire, ctx->stream) != 0) return ctx->fail(SCZ_ERR_NET, "net gather failed");
    return SCZ_OK;
}
int32_t CallbackNet::scatter(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) {
    if (is_leader()) upload += wire * (n_parties - 1);
    else download += wire;
    if (vt.scatter(vt.user, d_send, d_recv, bytes, wire, ctx->stream) != 0) return ctx->fail(SCZ_ERR_NET, "net scatter failed");
    return SCZ_OK;
}
int32_t CallbackNet::all_gather(Ctx *ctx, const void *d_send, void *d_recv, size_t bytes, size_t wire) {
    upload += wire * (n_parties - 1);
    download += wire * (n_parties - 1);
    if (vt.all_gather(vt.user, d_send, d_recv, bytes, wire, ctx->stream) != 0)
        return ctx->fail(SCZ_ERR_NET, "net all_gather failed");
    return SCZ_OK;
}
int32_t CallbackNet::gather_to(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire) {
    if (root == party_id) download += wire * (n_parties - 1);
    else upload += wire;
    if (!vt.gather_to || vt.gather_to(vt.user, root, d_send, d_recv, bytes, wire, ctx->stream) != 0)
        return ctx->fail(SCZ_ERR_NET, "net gather_to failed");
    return SCZ_OK;
}
int32_t CallbackNet::scatter_from(Ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire,
                                  bool *got) {
    if (root == party_id) upload += wire * (n_parties - 1);
    else download += wire;
    if (!vt.scatter_from || vt.scatter_from(vt.user, root, d_send, d_recv, bytes, wire, ctx->stream) != 0)
        return ctx->fail(SCZ_ERR_NET, "net scatter_from failed");
    if (got) *got = true;
    return SCZ_OK;
}
int32_t CallbackNet::sync(Ctx *ctx) {
    if (vt.sync && vt.sync(vt.user, ctx->stream) != 0) return ctx->fail(SCZ_ERR_NET, "net sync failed");
    return SCZ_OK;
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_ctx_create(int32_t device, uint32_t party_id, uint32_t n_parties, const scz_net_vtable *net, scz_ctx **out) {
    if (!out || n_parties == 0 || party_id >= n_parties) return SCZ_ERR_BAD_ARG;
    *out = nullptr;
    if (!net && party_id != 0) return SCZ_ERR_BAD_ARG;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        fprintf(stderr, "scz_ctx_create: no usable CUDA device %d (%s); libscz has no CPU path\n", device,
                e == cudaSuccess ? "index out of range" : cudaGetErrorString(e));
        return SCZ_ERR_CUDA;
    }
    scz_ctx *h = new scz_ctx();
    Ctx *c = &h->c;
    c->device = device;
    DeviceGuard dg(h);   // the caller's current device is restored on return
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return SCZ_ERR_CUDA;
    }
    c->own_stream = true;
    if (cudaMalloc(&c->d_status, sizeof(uint32_t)) != cudaSuccess || cudaMemset(c->d_status, 0, sizeof(uint32_t)) != cudaSuccess) {
        cudaStreamDestroy(c->stream);
        delete h;
        return SCZ_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    {   // keep freed temporaries cached in the pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            // reuse decided by stream order and event dependencies only, never by what the GPU happens to have finished:
            // with the opportunistic policy on, the pool kept growing into the 2nd and 3rd proof of a process depending on
            // how far the host ran ahead (4480 -> 4736 -> 4992 MiB at 2^20, +7.6 ms on the proof that grew it); without it
            // the pool is final after the first proof (4608 MiB) -- tools/pool_probe.py, profiles/r2_pool_probe.txt
            int off = 0;
            if (!getenv("SCZ_POOL_OPPORTUNISTIC")) cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowOpportunistic, &off);
        }
    }
    if (net) {
        CallbackNet *n = new CallbackNet();
        n->vt = *net;
        c->net = n;
    } else {
        c->net = new LeaderSimNet();
    }
    c->net->n_parties = n_parties;
    c->net->party_id = party_id;
    if (device < 64) g_live_ctx[device]++;
    if (const char *e = getenv("SCZ_MSM_AFFINE"))   // A/B switch for measurements: 0 = XYZZ only, 1 = always batched-affine
        c->msm_affine_mode = e[0] == '0' ? 2 : (e[0] == '1' ? 1 : 0);
    if (const char *e = getenv("SCZ_MSM_AFFINE_LEVELS")) c->msm_affine_levels = (uint32_t)atoi(e);
    *out = h;
    return SCZ_OK;
}
void scz_ctx_destroy(scz_ctx *h) {
    if (!h) return;
    Ctx *c = &h->c;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->d_status) cudaFree(c->d_status);
    if (c->stage[0].p) cudaFreeHost(c->stage[0].p);   // one block, sliced into the slots
    for (auto &st : c->stage)
        if (st.ev) cudaEventDestroy(st.ev);
    c->msm_affine_ws.reset();
    if (c->device < 64) g_live_ctx[c->device]--;
    c->prof_clear();
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (c->phase_mark) cudaEventDestroy(c->phase_mark);
    if (c->msm_stream) {
        cudaStreamSynchronize(c->msm_stream);
        cudaStreamDestroy(c->msm_stream);
        cudaEventDestroy(c->msm_fork);
        cudaEventDestroy(c->msm_join);
    }
    delete c->net;
    delete h;
}
const char *scz_last_error(const scz_ctx *h) { return h ? h->c.err.c_str() : "null ctx"; }
int32_t scz_ctx_set_stream(scz_ctx *h, void *s) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    SCZ_CUDA(c, cudaSetDevice(c->device));
    SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    return SCZ_OK;
}
int32_t scz_ctx_own_stream(scz_ctx *h) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (c->own_stream) return SCZ_OK;
    SCZ_CUDA(c, cudaSetDevice(c->device));
    SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    SCZ_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    return SCZ_OK;
}
int32_t scz_ctx_sync(scz_ctx *h) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    SCZ_CUDA(&h->c, cudaStreamSynchronize(h->c.stream));
    return SCZ_OK;
}
int32_t scz_prof_enable(scz_ctx *h, int32_t on) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    SCZ_CUDA(&h->c, cudaStreamSynchronize(h->c.stream));
    h->c.prof_clear();
    h->c.prof = on != 0;
    return SCZ_OK;
}
int32_t scz_prof_reserve(scz_ctx *h, uint64_t events) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (events > (1u << 22)) return c->fail(SCZ_ERR_BAD_ARG, "prof_reserve: %llu events", (unsigned long long)events);
    // every new event is also recorded once here: whatever the driver sets up lazily behind a timing event on its first
    // record happens now, not inside the region the caller is about to time
    bool fresh = false;
    while (c->prof_pool.size() < events) {
        cudaEvent_t e;
        SCZ_CUDA(c, cudaEventCreate(&e));
        SCZ_CUDA(c, cudaEventRecord(e, c->stream));
        c->prof_pool.push_back(e);
        fresh = true;
    }
    if (fresh) SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    return SCZ_OK;
}
int32_t scz_prof_read(scz_ctx *h, int32_t kernel_class, double *ms_total, uint64_t *brackets) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    double ms = 0;
    uint64_t n = 0;
    for (const Ctx::ProfRec &r : c->prof_recs) {
        if (r.id != kernel_class) continue;
        float t = 0;
        SCZ_CUDA(c, cudaEventElapsedTime(&t, r.a, r.b));
        ms += t;
        n++;
    }
    if (ms_total) *ms_total = ms;
    if (brackets) *brackets = n;
    return SCZ_OK;
}
uint64_t scz_ctx_launch_count(const scz_ctx *h) { return h ? h->c.launches : 0; }
int32_t scz_ctx_take_status(scz_ctx *h, uint32_t *bits) {
    scz::DeviceGuard dg__(h);
    if (!h || !bits) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    uint32_t v = 0;
    SCZ_CUDA(c, cudaMemcpyAsync(&v, c->d_status, sizeof v, cudaMemcpyDeviceToHost, c->stream));
    SCZ_CUDA(c, cudaMemsetAsync(c->d_status, 0, sizeof v, c->stream));
    SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    *bits = v;
    return SCZ_OK;
}
int32_t scz_ctx_status_snapshot_dev(scz_ctx *h, void *d_bits_out) {
    scz::DeviceGuard dg__(h);
    if (!h || !d_bits_out) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    SCZ_CUDA(c, cudaMemcpyAsync(d_bits_out, c->d_status, sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    SCZ_CUDA(c, cudaMemsetAsync(c->d_status, 0, sizeof(uint32_t), c->stream));
    return SCZ_OK;
}
int32_t scz_ctx_stream_wait_protocol_phase(scz_ctx *h, void *stream) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (c->phase_mark) SCZ_CUDA(c, cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(stream), c->phase_mark, 0));
    return SCZ_OK;
}
int32_t scz_ctx_get_comm(const scz_ctx *h, uint64_t *up, uint64_t *down) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (up) *up = h->c.net->upload;
    if (down) *down = h->c.net->download;
    return SCZ_OK;
}
int32_t scz_dev_alloc(scz_ctx *h, size_t bytes, void **p) {
    scz::DeviceGuard dg__(h);
    if (!h || !p) return SCZ_ERR_BAD_ARG;
    SCZ_CUDA(&h->c, cudaSetDevice(h->c.device));
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return h->c.fail(SCZ_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return SCZ_OK;
}
int32_t scz_dev_free(scz_ctx *h, void *p) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    SCZ_CUDA(&h->c, cudaStreamSynchronize(h->c.stream));
    SCZ_CUDA(&h->c, cudaFree(p));
    return SCZ_OK;
}
int32_t scz_h2d(scz_ctx *h, void *d, const void *s, size_t bytes) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    SCZ_CUDA(&h->c, cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->c.stream));
    SCZ_CUDA(&h->c, cudaStreamSynchronize(h->c.stream));
    return SCZ_OK;
}
int32_t scz_d2h(scz_ctx *h, void *d, const void *s, size_t bytes) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    SCZ_CUDA(&h->c, cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, h->c.stream));
    SCZ_CUDA(&h->c, cudaStreamSynchronize(h->c.stream));
    return SCZ_OK;
}
int32_t scz_host_alloc(scz_ctx *h, size_t bytes, void **p) {
    scz::DeviceGuard dg__(h);
    if (!h || !p) return SCZ_ERR_BAD_ARG;
    SCZ_CUDA(&h->c, cudaMallocHost(p, bytes ? bytes : 1));
    return SCZ_OK;
}
int32_t scz_host_free(scz_ctx *h, void *p) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    SCZ_CUDA(&h->c, cudaFreeHost(p));
    return SCZ_OK;
}

}   // extern "C"
