/*
 * ORACLE (test infrastructure): BLS12-381 field constants and vector helpers.
 * Moduli are the public BLS12-381 parameters (ark-bls12-381 0.4.0,
 * Cargo.lock:107-108); Montgomery constants are DERIVED from them at init so a
 * typo cannot hide in a hard-coded R or INV.
 */
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>
#include "oracle.h"

/* r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001 */
fr_t fr_MOD = {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}};
fr_t fr_R1, fr_R2;
uint64_t fr_INV;
/* p = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab */
fq_t fq_MOD = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
fq_t fq_R1, fq_R2;
uint64_t fq_INV;

void g1_init_(void);
static int inited = 0;
void orc_init(void) {
    if (inited) return;
    fr_derive_constants();
    fq_derive_constants();
    g1_init_();
    inited = 1;
}

#define VEC2(name, T, op)                                                   \
    void name(const T *a, const T *b, T *r, size_t n) {                     \
        for (size_t i = 0; i < n; i++) op(&r[i], &a[i], &b[i]);             \
    }
#define VEC1(name, T, op)                                                   \
    void name(const T *a, T *r, size_t n) {                                 \
        for (size_t i = 0; i < n; i++) op(&r[i], &a[i]);                    \
    }
VEC2(orc_fr_mul_vec, fr_t, fr_mul)
VEC2(orc_fr_add_vec, fr_t, fr_add)
VEC2(orc_fr_sub_vec, fr_t, fr_sub)
VEC1(orc_fr_to_canon_vec, fr_t, fr_to_canon)
VEC1(orc_fr_from_canon_vec, fr_t, fr_from_canon)
VEC2(orc_fq_mul_vec, fq_t, fq_mul)
VEC2(orc_fq_add_vec, fq_t, fq_add)
VEC2(orc_fq_sub_vec, fq_t, fq_sub)
VEC1(orc_fq_to_canon_vec, fq_t, fq_to_canon)
VEC1(orc_fq_from_canon_vec, fq_t, fq_from_canon)
void orc_fr_inv_vec(const fr_t *a, fr_t *r, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (!fr_inv(&r[i], &a[i])) fr_set_zero(&r[i]);
}

/* ---- pthread parallel-for ---- */
typedef struct { atomic_size_t next; size_t count; void (*body)(void *, size_t); void *ctx; } pf_t;
static void *pf_worker(void *arg) {
    pf_t *p = arg;
    for (;;) {
        size_t i = atomic_fetch_add(&p->next, 1);
        if (i >= p->count) break;
        p->body(p->ctx, i);
    }
    return NULL;
}
int orc_hw_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}
void orc_parallel_for(int threads, size_t count, void (*body)(void *, size_t), void *ctx) {
    if (threads > 256) threads = 256;
    if (threads <= 1 || count <= 1) {
        for (size_t i = 0; i < count; i++) body(ctx, i);
        return;
    }
    pf_t p;
    atomic_init(&p.next, 0);
    p.count = count;
    p.body = body;
    p.ctx = ctx;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < threads - 1; t++)
        if (pthread_create(&th[started], NULL, pf_worker, &p) == 0) started++;
    pf_worker(&p);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}
