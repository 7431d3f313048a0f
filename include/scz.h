/*
 * scz.h -- C ABI of libscz.so, the B200 (sm_100a) implementation of the
 * dist-primitive hot path of LBruyne/Scalable-Collaborative-zkSNARK.
 *
 * The reference has no FFI: its hot path is generic Rust over arkworks types.
 * Each entry point below names the reference function (file:line, relative to
 * the reference tree) whose CPU work it replaces; INTEGRATION.md shows the
 * `extern "C"` block and the thin Rust shim (`dist-primitive-gpu`) that keeps
 * the reference's own signatures on top of these calls.
 *
 * Data layout (identical to arkworks' in-memory layout, so Rust slices are
 * passed zero-copy):
 *   Fr          32 B   4 x u64 little-endian limbs, Montgomery form, R = 2^256
 *   Fq          48 B   6 x u64 little-endian limbs, Montgomery form, R = 2^384
 *   G1 affine   96 B   x | y (Fq each); the point at infinity is x = y = 0
 *                      (host entry points also take an optional byte mask that
 *                      mirrors ark-ec Affine::infinity)
 *   G1 Jacobian 144 B  X | Y | Z = ark-ec short_weierstrass::Projective;
 *                      identity has Z = 0
 *   triple      96 B   (Fr, Fr, Fr), one sumcheck round message
 *
 * Conventions: every function returns SCZ_OK (0) or a negative SCZ_ERR_*;
 * scz_last_error(ctx) gives the text.  The caller owns all host buffers; the
 * library never keeps a host pointer past return.  `_dev` entry points take
 * DEVICE pointers (from scz_dev_alloc or any CUDA allocation on ctx's device)
 * and run asynchronously on the ctx stream; the others take HOST pointers and
 * return after the result is in host memory.  A ctx is not thread-safe; use one
 * ctx per party / per host thread.  There is no CPU fallback anywhere: without
 * a CUDA device scz_ctx_create fails with SCZ_ERR_CUDA.
 */
#ifndef SCZ_H
#define SCZ_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCZ_OK 0
#define SCZ_ERR_BAD_ARG (-1)
#define SCZ_ERR_LEN_MISMATCH (-2) /* ark-ec msm Err(min len) -> the reference unwrap()s: dmsm.rs:23 */
#define SCZ_ERR_NOT_POW2 (-3)     /* dpoly_comm.rs:240,255 asserts */
#define SCZ_ERR_LEVEL_OOB (-4)    /* dpoly_comm.rs:239,254 asserts */
#define SCZ_ERR_CUDA (-5)
#define SCZ_ERR_NET (-6)          /* MPCNetError (mpc-net/src/lib.rs:14-26) */
#define SCZ_ERR_NOMEM (-7)

#define SCZ_FR_BYTES 32
#define SCZ_G1_AFFINE_BYTES 96
#define SCZ_G1_JAC_BYTES 144
#define SCZ_TRIPLE_BYTES 96
#define SCZ_G2_AFFINE_BYTES 192 /* x | y, each Fq2 = c0 | c1 (ark-ff Fp2); the identity is all zero */
#define SCZ_G2_JAC_BYTES 288    /* X | Y | Z = ark-ec Projective<g2::Config> */

typedef struct scz_ctx scz_ctx;
typedef struct scz_pp scz_pp;   /* PackedSharingParams<Fr>, secret-sharing/src/pss.rs:17-33 */
typedef struct scz_srs scz_srs; /* PolynomialCommitment<Bls12_381>, dpoly_comm.rs:30-34 (G1 part) */

/* ---- network seam: MPCSerializeNet (dist-primitive/src/utils/serializing_net.rs) ------------
 * All payloads are DEVICE buffers in device-native layout (no serialisation);
 * `wire_bytes` is the size the reference would have put on the wire
 * (ark-serialize compressed) and feeds the get_comm()-compatible counters.
 * Return 0 on success.  A NULL vtable selects the built-in leader simulator
 * (the reference's build without feature `comm`, serializing_net.rs:144-264). */
typedef struct scz_net_vtable {
    void *user;
    /* worker_send_or_leader_receive_element (:11-39): every party contributes `bytes` from d_send;
     * on the leader d_recv (n_parties * bytes, party-major) is filled. */
    int32_t (*gather)(void *user, const void *d_send, void *d_recv, size_t bytes, size_t wire_bytes, void *stream);
    /* worker_receive_or_leader_send_element (:76-96): leader's d_send holds n_parties * bytes; every party
     * receives its slice in d_recv. */
    int32_t (*scatter)(void *user, const void *d_send, void *d_recv, size_t bytes, size_t wire_bytes, void *stream);
    /* the N hub rounds of hyperplonk/src/dhyperplonk.rs:271-294 as one exchange: every party sends the same
     * `bytes` to all; d_recv (n_parties * bytes) is ordered by sender. */
    int32_t (*all_gather)(void *user, const void *d_send, void *d_recv, size_t bytes, size_t wire_bytes, void *stream);
    /* MPCNet::sync (mpc-net/src/lib.rs:275-286) */
    int32_t (*sync)(void *user, void *stream);
    /* the "dynamic" variants with a movable hub (serializing_net.rs:41-74, 98-126; used by c_acc_product_and_share):
     * everybody sends `bytes` to party `root`, whose d_recv (n_parties * bytes, party-major) is filled /
     * party `root`'s d_send holds n_parties * bytes, party j receives slice j.  May be NULL when cpermcheck is not used. */
    int32_t (*gather_to)(void *user, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire_bytes,
                         void *stream);
    int32_t (*scatter_from)(void *user, uint32_t root, const void *d_send, void *d_recv, size_t bytes, size_t wire_bytes,
                            void *stream);
} scz_net_vtable;

/* ---- context ---------------------------------------------------------------------------- */
/* party_id / n_parties mirror MPCNet::party_id / n_parties (mpc-net/src/lib.rs:37-41); party 0 is the leader.
 * net == NULL: leader simulator, party_id must be 0. */
int32_t scz_ctx_create(int32_t device, uint32_t party_id, uint32_t n_parties, const scz_net_vtable *net,
                       scz_ctx **out);
void scz_ctx_destroy(scz_ctx *ctx);
const char *scz_last_error(const scz_ctx *ctx);
/* run on a caller-owned CUDA stream (cudaStream_t, used as given: NULL is the legacy default stream) */
int32_t scz_ctx_set_stream(scz_ctx *ctx, void *cuda_stream);
/* go back to the ctx's private non-blocking stream (the state after scz_ctx_create) */
int32_t scz_ctx_own_stream(scz_ctx *ctx);
int32_t scz_ctx_sync(scz_ctx *ctx);
/* kernels launched so far by this ctx */
uint64_t scz_ctx_launch_count(const scz_ctx *ctx);
/* Per-kernel-class device timing for bench.py's roofline: when enabled, every launch group of a class is
 * bracketed by CUDA events on the ctx stream (the reference's start_timer!/end_timer! labels,
 * mpc-net/src/utils/timer.rs:25-197, are the host-side analogue).  scz_prof_read synchronises the stream
 * and returns the summed device time and the number of brackets of that class since scz_prof_enable. */
#define SCZ_K_MSM_SORT 0        /* digit recode, histogram, scan, scatter */
#define SCZ_K_MSM_ACCUMULATE 1  /* Pippenger bucket accumulation (the dominant stage): batched-affine levels + XYZZ, or XYZZ alone */
#define SCZ_K_MSM_FIXUP 2       /* buckets cut by chunk boundaries */
#define SCZ_K_MSM_REDUCE 3      /* bucket reduction: fan-in-8 tree (k_msm_tree) */
#define SCZ_K_MSM_FINISH 4      /* window recombination, Horner chain */
#define SCZ_K_PSS 5             /* PSS maps of the leader closures */
#define SCZ_K_SUMCHECK 6        /* fused fold + product-sum rounds */
#define SCZ_K_OPEN_FOLD 7       /* PST open: quotient + fold rounds */
#define SCZ_K_ACC_PRODUCT 8     /* product tree */
#define SCZ_K_POINTWISE 9       /* element-wise Fr maps, batch inversion */
int32_t scz_prof_enable(scz_ctx *ctx, int32_t on);
int32_t scz_prof_read(scz_ctx *ctx, int32_t kernel_class, double *ms_total, uint64_t *brackets);
/* Creates CUDA events ahead of time until the ctx's pool holds `events` of them (two per bracket): a measurement loop that
 * profiles K calls back to back reserves 2 * K * brackets-per-call first, so that no event is created inside its timed region */
int32_t scz_prof_reserve(scz_ctx *ctx, uint64_t events);
/* MPCNet::get_comm (mpc-net/src/lib.rs:59): (upload, download) in the reference's serialised bytes */
int32_t scz_ctx_get_comm(const scz_ctx *ctx, uint64_t *upload, uint64_t *download);
/* Sticky status bits of work already EXECUTED on the ctx stream; synchronises the stream, returns and clears them.
 * SCZ_STATUS_DIV_BY_ZERO: a field division (`num / den`, hyperplonk/src/dhyperplonk.rs:338-339) met den = 0 --
 * arkworks panics there ("division by zero" in Field::div); the kernel writes 0 for that element, sets this bit,
 * and the Rust shim panics when it sees it. */
#define SCZ_STATUS_DIV_BY_ZERO 1u
int32_t scz_ctx_take_status(scz_ctx *ctx, uint32_t *bits);
/* The same without synchronising: a copy of the bits (as of this point of the ctx stream) is written to the device word
 * d_bits_out and the sticky word is cleared, both in stream order.  For hosts that read a proof back asynchronously
 * (several proofs in flight): the snapshot travels with the proof's device -> host copy. */
int32_t scz_ctx_status_snapshot_dev(scz_ctx *ctx, void *d_bits_out);
/* Makes `stream` (a cudaStream_t of the host's) wait until the PROTOCOL PHASE of the prover call most recently enqueued on
 * this ctx has executed -- the ~1 100 short launches of dhyperplonk.rs:196-554 before the MSM sequences and leader rounds
 * take over (no-op before the first prover call).  A bulk host -> device copy (the next proof's tables) queued behind it
 * runs under the MSM phase; started at a proof boundary it shares PCIe with the command fetches of those short launches
 * and costs 6 ms per proof (measured, tools/e2e_probe.py). */
int32_t scz_ctx_stream_wait_protocol_phase(scz_ctx *ctx, void *stream);

/* ---- native data plane: NCCL over NVLink in place of the reference's TCP star (mpc-net/src/multi.rs:98-266) ----
 * One process per GPU ("rank"), `parties_per_rank` MPC parties hosted by each rank (1 on an 8-GPU box at l = 1; 8 l / W
 * on W GPUs), party id = rank * parties_per_rank + local index, party 0 = the leader (MPCNet::is_leader).  A hub owns
 * the rank's ncclComm_t; every hosted party gets its own ctx (one host thread per party, like the reference's one task
 * per party, multi.rs:345-348).  A star round (worker_send_or_leader_receive / worker_receive_or_leader_send,
 * mpc-net/src/lib.rs:64-256) is ONE grouped ncclSend / ncclRecv per rank on the concatenated payloads of its parties,
 * issued on the ctx stream of the rank's local party 0 -- no host copy, no serialisation, no Python; the hub rounds of
 * dhyperplonk.rs:271-294 are one ncclAllGather.  Local parties meet at a host barrier and are fenced by CUDA events.
 * The unique id is NCCL's (128 bytes): rank 0 creates it and the host hands it to the other ranks by whatever
 * channel it has (the reference's address file, MPI, torch.distributed ...). */
#define SCZ_NCCL_UID_BYTES 128
typedef struct scz_nccl_hub scz_nccl_hub;
int32_t scz_nccl_unique_id(void *uid128);
int32_t scz_nccl_hub_create(int32_t device, uint32_t rank, uint32_t nranks, uint32_t parties_per_rank, const void *uid128,
                            scz_nccl_hub **out);
void scz_nccl_hub_destroy(scz_nccl_hub *hub);
/* unblocks every party waiting at the hub's host barrier with SCZ_ERR_NET (a party thread failed) */
void scz_nccl_hub_abort(scz_nccl_hub *hub);
/* NCCL collectives issued so far by this rank: [gather, scatter, all_gather, sync] */
int32_t scz_nccl_hub_calls(const scz_nccl_hub *hub, uint64_t out4[4]);
/* the ctx of local party `local_index` (MPCNetConnection::init_from_path + listen + connect_to_all, multi.rs:109-266) */
int32_t scz_ctx_create_on_hub(scz_nccl_hub *hub, uint32_t local_index, scz_ctx **out);
/* one party per rank: hub + ctx in one call (the hub dies with the ctx) */
int32_t scz_ctx_create_nccl(int32_t device, uint32_t rank, uint32_t nranks, const void *uid128, scz_ctx **out);

/* The star collectives themselves on DEVICE buffers, for hosts that move their own elements (MPCNet's
 * worker_send_or_leader_receive :64, worker_receive_or_leader_send :165 and their dynamic_ variants :111, :211, sync
 * :275): whatever net the ctx was created with (leader simulator, callbacks, NCCL hub).  Asynchronous on the ctx stream.
 * gather: party `root` receives n_parties * bytes (party-major) in d_recv; scatter: party `root`'s d_send holds
 * n_parties * bytes and party j receives slice j in d_recv.  Byte counters advance by `bytes` per message. */
int32_t scz_net_gather(scz_ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes);
int32_t scz_net_scatter(scz_ctx *ctx, uint32_t root, const void *d_send, void *d_recv, size_t bytes);
int32_t scz_net_all_gather(scz_ctx *ctx, const void *d_send, void *d_recv, size_t bytes);
int32_t scz_net_sync(scz_ctx *ctx);
/* MPCNet::send_to / recv_from (mpc-net/src/lib.rs:55-61) as ncclSend / ncclRecv of `bytes` on the ctx stream: NCCL ctxs
 * with one party per rank only (the peer is the rank); the two sides must agree on `bytes` (the shim sends an 8-byte
 * length first, like the reference's length-delimited frames, mpc-net/src/multi.rs:29-35) */
int32_t scz_net_send(scz_ctx *ctx, uint32_t peer, const void *d_buf, size_t bytes);
int32_t scz_net_recv(scz_ctx *ctx, uint32_t peer, void *d_buf, size_t bytes);

/* ---- device memory ---------------------------------------------------------------------- */
int32_t scz_dev_alloc(scz_ctx *ctx, size_t bytes, void **d_ptr);
int32_t scz_dev_free(scz_ctx *ctx, void *d_ptr);
int32_t scz_h2d(scz_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int32_t scz_d2h(scz_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
/* pinned host memory for callers that want full-rate PCIe copies */
int32_t scz_host_alloc(scz_ctx *ctx, size_t bytes, void **h_ptr);
int32_t scz_host_free(scz_ctx *ctx, void *h_ptr);
/* fold an ark-ec `infinity` byte mask into the x = y = 0 encoding, in place on the device */
int32_t scz_g1_apply_inf_mask_dev(scz_ctx *ctx, void *d_bases, const uint8_t *d_mask, size_t n);

/* ---- element-wise kernels (unit-level parity with ark-ff / ark-ec) ------------------------ */
/* op: 0 add, 1 sub, 2 mul (ark-ff Fp Add/Sub/Mul) */
int32_t scz_fr_vec_op_dev(scz_ctx *ctx, int32_t op, const void *d_a, const void *d_b, void *d_out, size_t n);
int32_t scz_fq_vec_op_dev(scz_ctx *ctx, int32_t op, const void *d_a, const void *d_b, void *d_out, size_t n);
/* out[i] = a[i]^-1 (0 -> 0); Field::inverse */
int32_t scz_fr_inv_dev(scz_ctx *ctx, const void *d_a, void *d_out, size_t n);
/* Montgomery <-> canonical 32-byte little-endian integers (ark-ff into_bigint / from_bigint) */
int32_t scz_fr_to_canonical_dev(scz_ctx *ctx, const void *d_a, void *d_out, size_t n);
int32_t scz_fr_from_canonical_dev(scz_ctx *ctx, const void *d_a, void *d_out, size_t n);
/* out[i] = acc[i] (Jacobian) + (negate[i] ? -p[i] : p[i]) (affine); Projective += Affine */
int32_t scz_g1_add_affine_dev(scz_ctx *ctx, const void *d_acc_jac, const void *d_affine, const uint8_t *d_negate,
                              void *d_out_jac, size_t n);
/* op: 0 out = a + b, 1 out = 2a (b ignored); Projective Add / double_in_place */
int32_t scz_g1_vec_op_dev(scz_ctx *ctx, int32_t op, const void *d_a_jac, const void *d_b_jac, void *d_out_jac,
                          size_t n);
/* out[i] = k[i] * p[i]; Projective * Fr */
int32_t scz_g1_mul_fr_dev(scz_ctx *ctx, const void *d_jac, const void *d_k, void *d_out_jac, size_t n);
/* Projective::into_affine / CurveGroup::normalize_batch: Jacobian -> packed affine (infinity -> 0,0) */
int32_t scz_g1_to_affine_dev(scz_ctx *ctx, const void *d_jac, void *d_out_affine, size_t n);
/* ---- wire formats (ark-serialize compressed, what serializing_net.rs:17,25,50,60,88,95 put on the wire) ----
 * Fr: 32 B canonical little endian = scz_fr_to_canonical_dev; scz_fr_deserialize_dev also rejects values >= r
 * (status 1, like ark-ff's deserialize).  G1 (ark-bls12-381 0.4.0 = the Zcash / IETF encoding): 48 B, x big endian,
 * top bits of byte 0: compressed | infinity | y is the larger root.  Deserialisation validates like
 * deserialize_compressed (Validate::Yes): status 0 ok, 1 malformed / not on the curve, 2 not in the r-torsion. */
int32_t scz_g1_serialize_compressed_dev(scz_ctx *ctx, const void *d_jac, void *d_bytes48, size_t n);
int32_t scz_g1_deserialize_compressed_dev(scz_ctx *ctx, const void *d_bytes48, void *d_out_jac, uint8_t *d_status, size_t n);
int32_t scz_fr_deserialize_dev(scz_ctx *ctx, const void *d_bytes32, void *d_out, uint8_t *d_status, size_t n);
/* synthetic bases: out[i] = k[i] * G1 generator, packed affine (stands in for G1::rand, dpoly_comm.rs:214,229) */
int32_t scz_g1_generator_mul_dev(scz_ctx *ctx, const void *d_k, void *d_out_affine, size_t n);

/* ---- Fr tables: the local loops of dsumcheck.rs, dpoly_comm.rs, dacc_product.rs, mle.rs, dhyperplonk.rs ---- */
/* n = log2(len) rounds of the product sumcheck (dsumcheck.rs:37-85; the same loop at :167-219, :377-429):
 * round i splits the tables top-half / bottom-half, emits (sum f0 g0, sum f1 g1, sum (2f1-f0)(2g1-g0)) and folds with
 * challenge[i].  d_out_triples: n triples; d_last_fg: the two fully folded values (f, g).  Inputs are not modified. */
int32_t scz_sumcheck_product_rounds_dev(scz_ctx *ctx, const void *d_f, const void *d_g, size_t len,
                                        const void *d_challenge, void *d_out_triples, void *d_last_fg);
/* the fold rounds of a PST opening (dpoly_comm.rs:309-323 = :337-351 = :418-432): q_i = hi - lo, r = (1-u_i) lo + u_i hi.
 * d_q receives q_0 | q_1 | ... | q_{n-1} (len/2 + len/4 + ... + 1 = len - 1 entries), d_value the evaluation. */
int32_t scz_open_fold_dev(scz_ctx *ctx, const void *d_peval, size_t len, const void *d_point, void *d_q, void *d_value);
/* fix_variable (mle.rs:88-104): folds the top min(npoints, log2 len) variables; d_out gets len >> that many entries */
int32_t scz_fix_variable_dev(scz_ctx *ctx, const void *d_evals, size_t len, const void *d_points, size_t npoints,
                             void *d_out);
/* acc_product's table (dacc_product.rs:30-39 = :374-381): d_tree has 2m entries: x | products level by level | 0 */
int32_t scz_acc_product_dev(scz_ctx *ctx, const void *d_x, size_t m, void *d_tree);
/* point-wise maps of dhyperplonk.rs:233-238, 251-256, 326-339.  mode 0: a + b; 1: b - a; 2: a + k[0]*b + k[1]
 * (d_k: two Fr on the device); 3: a / b with ONE shared inversion per call (b = 0 -> 0 and SCZ_STATUS_DIV_BY_ZERO is raised:
 * arkworks panics there) */
int32_t scz_fr_pointwise_dev(scz_ctx *ctx, int32_t mode, const void *d_a, const void *d_b, const void *d_k, void *d_out,
                             size_t n);
/* even[i] = in[2i], odd[i] = in[2i+1]: v(x,0) and v(x,1) of the product tree (dhyperplonk.rs:349-359) */
int32_t scz_fr_deinterleave_dev(scz_ctx *ctx, const void *d_in, size_t n_pairs, void *d_even, void *d_odd);

/* ---- MSM: ark-ec VariableBaseMSM::msm, call sites dmsm.rs:23, dpoly_comm.rs:242,274,457 ----- */
/* One launch sequence computes `batch` independent MSMs; out_jac holds batch Jacobian points. */
int32_t scz_msm_g1_batched_dev(scz_ctx *ctx, const void *const *d_bases, const void *const *d_scalars,
                               const size_t *lens, size_t batch, void *d_out_jac);
/* host buffers; inf_mask may be NULL; bases_len != scalars_len -> SCZ_ERR_LEN_MISMATCH */
int32_t scz_msm_g1(scz_ctx *ctx, const void *bases, const uint8_t *inf_mask, size_t bases_len, const void *scalars,
                   size_t scalars_len, void *out_jac);
/* window override for experiments: 0 = automatic */
int32_t scz_msm_set_window(scz_ctx *ctx, uint32_t c);
/* Bucket accumulation variant (csrc/msm_affine.cu).  mode 0: automatic -- batched-affine additions (one shared field
 * inversion per tree level) inside the aligned single-bucket blocks of big sorted entry streams, XYZZ mixed additions
 * for the rest; 1: always; 2: never.  levels: affine tree levels, 0 = automatic; slab_entries: entries processed per
 * pass, 0 = from the free device memory.  The group elements are the same in every mode (tests/test_gpu_msm.py). */
int32_t scz_msm_set_affine(scz_ctx *ctx, uint32_t mode, uint32_t levels, uint64_t slab_entries);
/* launch sequences of this ctx whose bucket accumulation took the batched-affine path */
int32_t scz_msm_affine_sequences(const scz_ctx *ctx, uint64_t *sequences);
/* statistics of the last MSM launch sequence: total (point, window) pairs = bucket additions, buckets, windows */
int32_t scz_msm_last_stats(const scz_ctx *ctx, uint64_t *bucket_adds, uint64_t *buckets, uint64_t *windows);
/* cumulative since scz_ctx_create: bucket additions, (base, scalar) pairs, launch sequences, MSMs (segments) */
int32_t scz_msm_cum_stats(const scz_ctx *ctx, uint64_t *bucket_adds, uint64_t *pairs, uint64_t *sequences,
                          uint64_t *segments);

/* ---- PSS: secret-sharing/src/pss.rs ------------------------------------------------------- */
int32_t scz_pp_new(scz_ctx *ctx, size_t l, scz_pp **out);          /* PackedSharingParams::new, pss.rs:38-65 */
void scz_pp_free(scz_pp *pp);
int32_t scz_pp_info(const scz_pp *pp, size_t *t, size_t *l, size_t *n);
/* kind: 0 = Fr (32 B), 1 = G1 Jacobian (144 B).  `batch` independent vectors, vector-major.
 * pack_from_public (pss.rs:69-99): in batch x len_in (len_in <= 2l, zero padded) -> out batch x n
 * pack_single      (pss.rs:103-113): in batch x 1 -> out batch x n
 * unpack           (pss.rs:117-149): in batch x n -> out batch x l
 * unpack2          (pss.rs:124-171): in batch x n -> out batch x l */
int32_t scz_pss_pack_from_public_dev(scz_ctx *ctx, const scz_pp *pp, int32_t kind, const void *d_in, size_t len_in,
                                     size_t batch, void *d_out);
int32_t scz_pss_pack_single_dev(scz_ctx *ctx, const scz_pp *pp, int32_t kind, const void *d_in, size_t batch,
                                void *d_out);
int32_t scz_pss_unpack_dev(scz_ctx *ctx, const scz_pp *pp, int32_t kind, const void *d_in, size_t batch, void *d_out);
int32_t scz_pss_unpack2_dev(scz_ctx *ctx, const scz_pp *pp, int32_t kind, const void *d_in, size_t batch,
                            void *d_out);

/* ---- d_msm: dist-primitive/src/dmsm.rs:9-43 ------------------------------------------------ */
/* local MSMs -> gather -> leader: unpack2, sum, pack_from_public -> scatter.  out: batch Jacobian points
 * (this party's packed shares of the batch results). */
int32_t scz_d_msm_dev(scz_ctx *ctx, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars,
                      const size_t *lens, size_t batch, void *d_out_jac);
/* The two halves of d_msm for hosts that run several parties in one process (fewer GPUs than parties):
 * the local half is scz_msm_g1_batched_dev (dmsm.rs:19-24); this is the leader closure (dmsm.rs:31-38) on
 * the gathered buffer, party-major [party][k], n_parties x batch Jacobian points in and out. */
int32_t scz_d_msm_leader_dev(scz_ctx *ctx, const scz_pp *pp, const void *d_gathered, size_t batch,
                             void *d_to_scatter);
int32_t scz_d_msm(scz_ctx *ctx, const scz_pp *pp, const void *const *bases, const size_t *bases_lens,
                  const void *const *scalars, const size_t *scalars_lens, size_t batch, void *out_jac);

/* ---- the same over G2 (`d_msm` is generic over `G: CurveGroup`, dmsm.rs:9-15; BASELINE names "d_msm over G1/G2"; the
 * reference itself only ever instantiates G1).  Affine bases 192 B, results Jacobian 288 B (SCZ_G2_*).  Same semantics as the
 * G1 entry points; a plain, untuned Pippenger that shares the G1 path's digit recoding and counting sort (csrc/msm_g2.cu). */
int32_t scz_msm_g2_batched_dev(scz_ctx *ctx, const void *const *d_bases, const void *const *d_scalars, const size_t *lens,
                               size_t batch, void *d_out_jac);
int32_t scz_msm_g2(scz_ctx *ctx, const void *bases, const uint8_t *inf_mask, size_t bases_len, const void *scalars,
                   size_t scalars_len, void *out_jac);
int32_t scz_d_msm_g2_dev(scz_ctx *ctx, const scz_pp *pp, const void *const *d_bases, const void *const *d_scalars,
                         const size_t *lens, size_t batch, void *d_out_jac);
int32_t scz_d_msm_g2_leader_dev(scz_ctx *ctx, const scz_pp *pp, const void *d_gathered, size_t batch, void *d_to_scatter);
/* unit operations on Jacobian G2 points for the parity tests: op 0 out = a + b, 1 out = 2 a, 2 out = k * a (d_b: Fr) */
int32_t scz_g2_vec_op_dev(scz_ctx *ctx, int32_t op, const void *d_a_jac, const void *d_b, void *d_out_jac, size_t n);

/* ---- re-sharing rounds -------------------------------------------------------------------------------- */
/* pss2ss (unpack.rs:72-97): one share in, Vec<F> of length l out (gather, leader unpack + pack_single, scatter) */
int32_t scz_pss2ss_dev(scz_ctx *ctx, const scz_pp *pp, const void *d_share, void *d_out);
/* degree_reduce (degree_reduce.rs:29-41): one share in, one share out (gather, unpack2, pack_from_public, scatter) */
int32_t scz_degree_reduce_dev(scz_ctx *ctx, const scz_pp *pp, const void *d_share, void *d_out);

/* ---- sumcheck family: dist-primitive/src/dsumcheck.rs ---------------------------------------------------- */
/* sumcheck_product (:28-90): n + 1 triples, the last is (0, f*g, 0) */
int32_t scz_sumcheck_product_dev(scz_ctx *ctx, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                                 void *d_out);
/* c_sumcheck_product (:148-285): n + log2(l) + 1 triples; phase 2 re-uses challenge[0..log2 l) like the reference (:230) */
int32_t scz_c_sumcheck_product_dev(scz_ctx *ctx, const scz_pp *pp, const void *d_f, const void *d_g, size_t len,
                                   const void *d_challenge, void *d_out);
/* d_sumcheck_product (:359-512): d_challenge holds n + log2(N) challenges; the leader gets *count = n + log2(N)
 * triples in d_out, every other party *count = 0 (the reference returns an empty Vec, :507-509) */
int32_t scz_d_sumcheck_product_dev(scz_ctx *ctx, const void *d_f, const void *d_g, size_t len, const void *d_challenge,
                                   void *d_out, size_t *count);
/* single-MLE variants (round message = (sum lo, sum hi), 64 B): sumcheck (:6-26) n + 1 pairs, the last is (0, f(r));
 * c_sumcheck (:92-146) n + log2(l) + 1 pairs; d_sumcheck (:287-357) leader n + log2(N) pairs, others *count = 0 */
int32_t scz_sumcheck_dev(scz_ctx *ctx, const void *d_f, size_t len, const void *d_challenge, void *d_out);
int32_t scz_c_sumcheck_dev(scz_ctx *ctx, const scz_pp *pp, const void *d_f, size_t len, const void *d_challenge,
                           void *d_out);
int32_t scz_d_sumcheck_dev(scz_ctx *ctx, const void *d_f, size_t len, const void *d_challenge, void *d_out, size_t *count);
/* d_acc_product (dacc_product.rs:365-414): d_subtree 2m entries; on the leader d_leader_tree 2N entries */
int32_t scz_d_acc_product_dev(scz_ctx *ctx, const void *d_x, size_t m, void *d_subtree, void *d_leader_tree);

/* ---- multilinear KZG: dist-primitive/src/dpoly_comm.rs --------------------------------------------------- */
/* powers_of_g (:30-34): level i = packed affine bases; _device_ borrows the caller's device arrays, _host_ uploads */
int32_t scz_srs_from_device_levels(scz_ctx *ctx, size_t levels, const void *const *d_levels, const size_t *lens,
                                   scz_srs **out);
int32_t scz_srs_from_host_levels(scz_ctx *ctx, size_t levels, const void *const *levels_host, const size_t *lens,
                                 scz_srs **out);
/* PolynomialCommitmentCub::new(g, _, s).mature() (dpoly_comm.rs:37-67, 141-150): the real SRS with trapdoor s[0..n),
 * levels 0..n of 2^i points, built on the device (set-up) */
int32_t scz_srs_new_dev(scz_ctx *ctx, const void *d_g_jac, const void *d_s, size_t n, scz_srs **out);
/* to_packed (dpoly_comm.rs:164-194): PSS-pack every level over G1 in chunks of l points; returns party `party`'s SRS
 * (levels of max(1, 2^i / l) share points) */
int32_t scz_srs_to_packed_dev(scz_ctx *ctx, const scz_srs *srs, const scz_pp *pp, uint32_t party, scz_srs **out);
/* copy a level (packed affine) into d_out (may be NULL to query the length) */
int32_t scz_srs_level_dev(scz_ctx *ctx, const scz_srs *srs, size_t level, void *d_out, size_t *len);
void scz_srs_free(scz_srs *srs);
/* Fixed-base tables for every level: the window multiples 2^(c w) * P_j next to the points, so that all windows of a
 * scalar share one bucket set (fewer bucket additions, no Horner recombination; see csrc/srs.cu).  Same results,
 * about 13 extra copies of the SRS in HBM.  MSMs of the commit / open family that cover a whole level use the table. */
int32_t scz_srs_precompute(scz_ctx *ctx, scz_srs *srs);
/* on = 0: this ctx ignores fixed-base tables (A/B measurements); default 1 */
int32_t scz_msm_use_precompute(scz_ctx *ctx, int32_t on);
int32_t scz_srs_info(const scz_srs *srs, size_t *levels);
/* commit / d_local_commit (:237-243, :269-275): len must be 2^level with level < #levels (SCZ_ERR_NOT_POW2 / LEVEL_OOB) */
int32_t scz_commit_dev(scz_ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, void *d_out_jac);
/* c_commit (:244-267): level = log2(len * l); one d_msm over the batch */
int32_t scz_c_commit_dev(scz_ctx *ctx, const scz_srs *srs, const scz_pp *pp, const void *const *d_pevals,
                         const size_t *lens, size_t batch, void *d_out_jac);
/* d_commit (:276-297): local commit, the leader sums the N commitments, every party receives the sum */
int32_t scz_d_commit_dev(scz_ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, void *d_out_jac);
/* open / d_local_open (:299-325, :327-353): value (1 Fr) + n proofs */
int32_t scz_open_dev(scz_ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point,
                     void *d_value, void *d_proofs_jac);
/* c_open (:401-464): value + n + log2(l) proofs; all n quotient commitments go out as one batched c_commit (:436) */
int32_t scz_c_open_dev(scz_ctx *ctx, const scz_srs *srs, const scz_pp *pp, const void *d_peval, size_t len,
                       const void *d_point, void *d_value, void *d_proofs_jac);
/* d_open (:355-398): point has log2(N) + n coordinates; leader: value + log2(N) root proofs ++ n summed proofs
 * (*count = log2(N) + n); others: value 0, *count = 0 (:387) */
int32_t scz_d_open_dev(scz_ctx *ctx, const scz_srs *srs, const void *d_peval, size_t len, const void *d_point,
                       size_t npoint, void *d_value, void *d_proofs_jac, size_t *count);

/* ---- the prover: hyperplonk/src/dhyperplonk.rs:159-571 ---------------------------------------------------------
 * scz_hp_pk mirrors the fields of PackedProvingParameters (dhyperplonk.rs:22-62) that `dhyperplonk` reads: DEVICE
 * pointers to Fr tables (lengths for circuit size 2^n, packing factor l, N parties; gc = 2^n):
 *   V 4gc/l; a_evals, b_evals, c_evals, I, S1, S2, eq gc/l; I_p, S1_p, S2_p gc/N; ssigma_p, sid_p, eq_r1_p, eq_r2_p 4gc/N;
 *   challenge n; challenge_r1, challenge_r2 n+2; alpha_beta 2 (alpha | beta).
 * The three vectors the reference draws from entropy inside the function (:188-190) are explicit inputs (an API
 * addition so that runs are reproducible): local_s_p 4gc/N, local_s 4gc/N/l, eq_leader 8l. */
typedef struct scz_hp_pk {
    const void *V, *a_evals, *b_evals, *c_evals, *I, *S1, *S2, *I_p, *S1_p, *S2_p, *ssigma_p, *sid_p, *eq, *eq_r1_p,
        *eq_r2_p, *challenge, *challenge_r1, *challenge_r2, *alpha_beta;
    const scz_srs *c_commitment; /* new_single(n + 2, pp), dhyperplonk.rs:100 */
    const scz_srs *d_commitment; /* new_random(n + 2, N), dhyperplonk.rs:101 */
    const void *local_s_p, *local_s, *eq_leader;
} scz_hp_pk;
/* One entry of the reference's return value (dhyperplonk.rs:567-570), in the order the reference pushes it into
 * its vector.  Offsets are element indices into the three output arenas. */
#define SCZ_HP_GATE_PROOF 0    /* gate_identity_proofs[i]: triples                                         */
#define SCZ_HP_GATE_COMMIT 1   /* gate_identity_commitments[i]: points[0] = commitment, then the proofs; value  */
#define SCZ_HP_WIRING_PROOF 2  /* wiring_proofs[i]: triples (empty on non-leaders for the d_ variants)        */
#define SCZ_HP_WIRING_COMMIT 3 /* wiring_commits[i]: one point                                               */
#define SCZ_HP_WIRING_OPEN 4   /* wiring_opens[i]: value + proofs (value 0, no proofs on non-leaders for d_open) */
typedef struct scz_hp_item {
    uint32_t kind;
    uint32_t triples_off, triples_cnt; /* into d_triples (96 B each) */
    uint32_t points_off, points_cnt;   /* into d_points  (144 B Jacobian each) */
    uint32_t value_off, value_cnt;     /* into d_values  (32 B each); value_cnt is 0 or 1 */
} scz_hp_item;
/* capacities (in elements / items) that always suffice for the outputs of scz_dhyperplonk_dev */
int32_t scz_dhyperplonk_sizes(size_t n, size_t l, size_t n_parties, size_t *triples, size_t *points, size_t *values,
                              size_t *items);
/* dhyperplonk (dhyperplonk.rs:159-571) from after net.sync() (:193) to the return.  Outputs go to three DEVICE arenas;
 * `items` (HOST array) describes them entry by entry, *n_items entries.  Asynchronous on the ctx stream except for
 * the host-side collectives of a real net. */
int32_t scz_dhyperplonk_dev(scz_ctx *ctx, size_t n, const scz_hp_pk *pk, const scz_pp *pp, void *d_triples,
                            size_t triples_cap, void *d_points, size_t points_cap, void *d_values, size_t values_cap,
                            scz_hp_item *items, size_t items_cap, size_t *n_items);
/* dhyperplonk_data_parallel (dhyperplonk.rs:573-960): the same schedule, but `s` of step 2.a is an input instead of an
 * exchange (:603): pk->local_s holds 4gc/l entries.  Same outputs layout. */
int32_t scz_dhyperplonk_data_parallel_dev(scz_ctx *ctx, size_t n, const scz_hp_pk *pk, const scz_pp *pp, void *d_triples,
                                          size_t triples_cap, void *d_points, size_t points_cap, void *d_values,
                                          size_t values_cap, scz_hp_item *items, size_t items_cap, size_t *n_items);
/* dpermcheck (dhyperplonk.rs:962-1247): the wiring identity (step 2) alone; only SCZ_HP_WIRING_* items come back.
 * The gate-side fields of pk (a/b/c_evals, I, S1, S2, eq, challenge, *_p selectors) are not read but must be non-null. */
int32_t scz_dpermcheck_dev(scz_ctx *ctx, size_t n, const scz_hp_pk *pk, const scz_pp *pp, void *d_triples, size_t triples_cap,
                           void *d_points, size_t points_cap, void *d_values, size_t values_cap, scz_hp_item *items,
                           size_t items_cap, size_t *n_items);

/* ---- local_hyperplonk (hyperplonk/src/hyperplonk.rs:15-160): the monolithic single-prover baseline ("Local HyperPlonk").
 * Plain DEVICE tables: a_evals, b_evals, c_evals, input, q1, q2, eq: 2^n entries; m, ssigma, sid, eq_p2: 4 * 2^n;
 * challenge n, challengep2 n + 2, alpha_beta 2 (num = m + alpha sid + beta, :106-115); commitment: levels 0 .. n+2.
 * Returns 6 gate proofs, 6 gate commit+open entries, 6 wiring proofs, 8 wiring commits, 8 wiring opens. */
typedef struct scz_local_pk {
    const void *m, *a_evals, *b_evals, *c_evals, *input, *q1, *q2, *ssigma, *sid, *eq, *eq_p2, *challenge, *challengep2,
        *alpha_beta;
    const scz_srs *commitment;
} scz_local_pk;
int32_t scz_local_hyperplonk_dev(scz_ctx *ctx, size_t n, const scz_local_pk *pk, void *d_triples, size_t triples_cap,
                                 void *d_points, size_t points_cap, void *d_values, size_t values_cap, scz_hp_item *items,
                                 size_t items_cap, size_t *n_items);

/* ---- the collaborative (PSS) permutation check, the paper's baseline: hyperplonk/src/dhyperplonk.rs:1249-1385 -----
 * c_acc_product_and_share (dacc_product.rs:66-292): masked product accumulation over N hub rounds with a moving hub.
 * shares / masks / unmask0..2 and the three outputs (v(x,0), v(x,1), v(1,x) shares) hold `len` entries each; len / N * l
 * must be a power of two >= N.  Needs the gather_to / scatter_from callbacks of a real net. */
int32_t scz_c_acc_product_and_share_dev(scz_ctx *ctx, const scz_pp *pp, const void *d_shares, const void *d_masks,
                                        const void *d_unmask0, const void *d_unmask1, const void *d_unmask2, size_t len,
                                        void *d_share0, void *d_share1, void *d_share2);
/* the fields of PackedProvingParameters that cpermcheck reads: DEVICE tables of 4 * 2^n / l entries (V, sid, ssigma,
 * eq_r1, mask, unmask0..2), challenge_r1 (n + 2), alpha_beta (2) and the collaborative SRS */
typedef struct scz_cperm_pk {
    const void *V, *sid, *ssigma, *eq_r1, *mask, *unmask0, *unmask1, *unmask2, *challenge_r1, *alpha_beta;
    const scz_srs *c_commitment;
} scz_cperm_pk;
/* cpermcheck: 6 SCZ_HP_WIRING_PROOF, 10 SCZ_HP_WIRING_COMMIT and 12 SCZ_HP_WIRING_OPEN items in the reference's push
 * order; arenas sized by scz_dhyperplonk_sizes */
int32_t scz_cpermcheck_dev(scz_ctx *ctx, size_t n, const scz_cperm_pk *pk, const scz_pp *pp, void *d_triples,
                           size_t triples_cap, void *d_points, size_t points_cap, void *d_values, size_t values_cap,
                           scz_hp_item *items, size_t items_cap, size_t *n_items);

#ifdef __cplusplus
}
#endif
#endif
