// Library-internal context: device, stream, grow-only workspace arena, error text.
// One scz_ctx per MPC party (the reference spawns one task per party,
// mpc-net/src/multi.rs:345-348); calls on a ctx are serialised by the caller.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <memory>
#include <string>
#include <vector>

#include "../../include/scz.h"

namespace scz {

struct Net;

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t msm_stream = nullptr;         // optional lowest-priority stream of the MSM launch sequences (msm.cu, flush_msm)
    cudaEvent_t msm_fork = nullptr, msm_join = nullptr;
    // recorded on `stream` where a prover's protocol phase ends and only its MSM sequences / leader rounds remain
    // (Deferred::run); scz_ctx_stream_wait_protocol_phase lets a host copy stream wait for it
    cudaEvent_t phase_mark = nullptr;
    int sm_count = 148;
    std::string err;
    // pinned host staging for small results
    char *pinned = nullptr;
    size_t pinned_cap = 0;
    Net *net = nullptr;
    uint64_t launches = 0;   // kernels launched by this ctx (bench.py reports it as gpu_launches)
    // cudaFuncSetAttribute is per device: every ctx raises the dynamic shared memory limits it needs once (no
    // process-wide flags -- a process may drive several devices and several party threads)
    bool attr_msm = false, attr_pss = false;
    uint32_t *d_status = nullptr;   // sticky SCZ_STATUS_* bits set by kernels (scz_ctx_take_status)
    uint32_t msm_window_override = 0;
    bool msm_no_precompute = false;   // ignore fixed-base tables (for A/B measurements)
    // batched-affine bucket accumulation (msm_affine.cu): 0 = automatic (big sequences with long bucket runs), 1 = always,
    // 2 = never; levels / slab entries: 0 = automatic
    uint32_t msm_affine_mode = 0, msm_affine_levels = 0;
    uint64_t msm_affine_slab = 0;
    uint64_t msm_cum_affine_sequences = 0;
    // the level arrays of the batched-affine accumulation: allocated once per ctx and kept (grow-only).  Per-sequence
    // stream-ordered allocations of tens of GB made the pool re-map memory when proofs are enqueued back to back
    // (254 instead of 122 ms per proof measured), so this workspace is NOT a DevTmp.
    std::shared_ptr<void> msm_affine_ws;
    uint64_t msm_affine_ws_limit = UINT64_MAX;   // slab entries above which an allocation failed on this device
    uint64_t msm_bucket_adds = 0, msm_buckets = 0, msm_windows = 0;   // statistics of the last MSM sequence
    uint64_t msm_cum_adds = 0, msm_cum_pairs = 0, msm_cum_sequences = 0, msm_cum_segments = 0;   // since ctx creation
    // optional per-kernel-class device timing (scz_prof_*): CUDA events recorded on `stream` around the launches
    struct ProfRec {
        int id;
        cudaEvent_t a, b;
    };
    bool prof = false;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    void prof_begin(int id);
    void prof_end();
    void prof_clear();

    // Small host -> device tables (MSM segment tables, closure job lists) go through a ring of pinned slots so that
    // the copy is truly asynchronous and the host can run ahead of the GPU; a slot is reused only after the copy
    // that last used it has executed.
    // 32 slots: a 2^20 proof stages ~10 tables, and the one of its second MSM sequence sits behind ~115 ms of bucket
    // kernels on the side stream -- with 8 slots the host met that copy's event again inside the same proof and stalled
    // until the first sequence had finished (host enqueue 115 ms instead of 12 ms per proof).  The pinned memory of all
    // slots is allocated in one piece on first use: cudaMallocHost synchronises the device, and slots allocated one by
    // one as the ring advanced stalled the first three or four proofs of a process (432 ms for the first timed proof of a
    // bench run with three warm-up proofs, 153 ms for the rest).
    static constexpr int STAGE_SLOTS = 32;
    static constexpr size_t STAGE_BYTES = 1 << 20;
    struct Stage {
        char *p = nullptr;
        cudaEvent_t ev = nullptr;
        bool busy = false;
    } stage[STAGE_SLOTS];
    int stage_next = 0;
    int32_t h2d_staged(void *d_dst, const void *h_src, size_t bytes);

    int32_t fail(int32_t code, const char *fmt, ...);
    int32_t cuda(cudaError_t e, const char *what);
    int32_t pinned_reserve(size_t bytes);
};

// Stream-ordered temporary from the device's CUDA memory pool (cudaMallocAsync): freed in
// stream order when the object dies, so kernels already queued keep their memory and
// steady-state calls never hit the allocator's slow path (release threshold = unlimited).
// `persistent`: a plain cudaMalloc that lives until the object dies (workspaces kept across calls).
struct DevTmp {
    Ctx *c;
    void *p = nullptr;
    bool persistent = false;
    explicit DevTmp(Ctx *ctx, bool keep = false) : c(ctx), persistent(keep) {}
    DevTmp(const DevTmp &) = delete;
    DevTmp &operator=(const DevTmp &) = delete;
    int32_t alloc(size_t bytes) {
        cudaError_t e = persistent ? cudaMalloc(&p, bytes ? bytes : 256) : cudaMallocAsync(&p, bytes ? bytes : 256, c->stream);
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            return c->fail(SCZ_ERR_NOMEM, "%s(%zu): %s", persistent ? "cudaMalloc" : "cudaMallocAsync", bytes, cudaGetErrorString(e));
        }
        return SCZ_OK;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    ~DevTmp() {
        if (!p) return;
        if (persistent) cudaFree(p);
        else cudaFreeAsync(p, c->stream);
    }
};
// live ctxs on a device (several parties may share one GPU: workspaces are sized accordingly)
int ctx_live_on_device(int device);

#define SCZ_CUDA(ctx, expr)                                   \
    do {                                                      \
        cudaError_t e__ = (expr);                             \
        if (e__ != cudaSuccess) return (ctx)->cuda(e__, #expr); \
    } while (0)
#define SCZ_TRY(expr)                  \
    do {                               \
        int32_t rc__ = (expr);         \
        if (rc__ != SCZ_OK) return rc__; \
    } while (0)
#define SCZ_LAUNCH_CHECK(ctx)                                         \
    do {                                                              \
        (ctx)->launches++;                                            \
        cudaError_t e__ = cudaGetLastError();                         \
        if (e__ != cudaSuccess) return (ctx)->cuda(e__, "kernel launch"); \
    } while (0)

// RAII bracket: times everything launched on ctx->stream inside the scope under kernel class `id`
struct ProfScope {
    Ctx *c;
    ProfScope(Ctx *ctx, int id) : c(ctx) {
        if (c->prof) c->prof_begin(id);
    }
    ~ProfScope() {
        if (c->prof) c->prof_end();
    }
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline uint32_t ceil_div_u32(size_t a, size_t b) { return (uint32_t)((a + b - 1) / b); }

}   // namespace scz

struct scz_ctx {
    scz::Ctx c;
};

namespace scz {
// Every ABI entry point runs on the ctx's device whatever the calling thread's current device is (the reference's
// party tasks hop OS threads, mpc-net/src/multi.rs:345-348), and leaves the caller's device as it found it.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const scz_ctx *h) {
        if (!h) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != h->c.device) {
            prev = cur;
            cudaSetDevice(h->c.device);
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
}   // namespace scz
