"""Python mirror of the reference's Rust interface for the hot path.

Names, argument meaning and error behaviour follow the reference
(`PackedSharingParams` secret-sharing/src/pss.rs:17-171, `d_msm`
dist-primitive/src/dmsm.rs:9-43, ...).  Arrays use arkworks' in-memory layout:

  Fr          (n, 4)  uint64   Montgomery limbs
  G1 affine   (n, 12) uint64   x | y, infinity = all zero
  G1 Jacobian (n, 18) uint64   X | Y | Z

Every function accepts either numpy arrays (HOST path: the C ABI copies in and
out, like a call from the Rust shim) or torch CUDA tensors of dtype int64 with the
same shapes (DEVICE path: tables stay resident in HBM between calls).
"""
import ctypes as C

import numpy as np
import torch

from .binding import NetVTable, SczError, lib

FR_LIMBS, AFF_LIMBS, JAC_LIMBS = 4, 12, 18


def _is_dev(x):
    return isinstance(x, torch.Tensor)


def _host(a, cols):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, cols)


def _dev(t, cols):
    assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous(), "device operands: contiguous int64 CUDA tensors"
    return t.reshape(-1, cols)


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


class Context:
    """One MPC party's handle (scz_ctx).  `net=None` selects the reference's leader
    simulator (build without feature `comm`, serializing_net.rs:144-264)."""

    def __init__(self, device=0, party_id=0, n_parties=8, net=None):
        if not torch.cuda.is_available():
            raise SczError(-5, "no CUDA device: scz-b200 has no CPU path")
        self.L = lib()
        self.device = torch.device("cuda", device)
        self.net = net
        self._vt = None
        h = C.c_void_p()
        if net is not None and hasattr(net, "native_create"):
            # libscz's own NCCL hub (csrc/nccl_net.cu): the collectives never come back to Python
            assert (party_id, n_parties) == (net.rank, net.n_parties), "party id / count are the hub's"
            rc = net.native_create(self.L, h)
        else:
            if net is not None:
                self._vt = net.vtable()
            rc = self.L.scz_ctx_create(C.c_int32(device), C.c_uint32(party_id), C.c_uint32(n_parties),
                                       C.byref(self._vt) if self._vt is not None else None, C.byref(h))
        if rc != 0:
            raise SczError(rc, "scz_ctx_create failed")
        self.h = h
        self.party_id, self.n_parties = party_id, n_parties
        self.use_torch_stream()

    def use_torch_stream(self):
        """run on torch's current stream so torch.cuda.Event timing and tensors order with our kernels"""
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream()
        self.check(self.L.scz_ctx_set_stream(self.h, C.c_void_p(self.stream.cuda_stream)))

    def check(self, rc):
        if rc != 0:
            raise SczError(rc, self.L.scz_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.scz_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(self.L.scz_ctx_sync(self.h))

    def stream_wait_protocol_phase(self, stream):
        """`stream` (torch.cuda.Stream) waits until the protocol phase of the prover call last enqueued on this ctx has
        run (scz_ctx_stream_wait_protocol_phase): queue the next proof's host -> device copy behind it"""
        self.check(self.L.scz_ctx_stream_wait_protocol_phase(self.h, C.c_void_p(stream.cuda_stream)))

    def take_status(self):
        """sticky SCZ_STATUS_* bits of the work executed so far (synchronises); bit 0: a division met a zero denominator"""
        bits = C.c_uint32()
        self.check(self.L.scz_ctx_take_status(self.h, C.byref(bits)))
        return bits.value

    def raise_on_status(self):
        """what the Rust shim does after a prover call: arkworks panics on `a / 0` (dhyperplonk.rs:338-339)"""
        if self.take_status() & 1:
            raise ZeroDivisionError("field division by zero (arkworks panics here: hyperplonk/src/dhyperplonk.rs:338-339)")

    @property
    def launches(self):
        return int(self.L.scz_ctx_launch_count(self.h))

    KERNEL_CLASSES = {"msm_sort": 0, "msm_accumulate": 1, "msm_fixup": 2, "msm_reduce": 3, "msm_finish": 4, "pss": 5,
                      "sumcheck": 6, "open_fold": 7, "acc_product": 8, "pointwise": 9}

    def prof_enable(self, on=True):
        """bracket every kernel class with CUDA events on the ctx stream (scz_prof_enable); clears old records"""
        self.check(self.L.scz_prof_enable(self.h, C.c_int32(1 if on else 0)))

    def prof_reserve(self, events):
        """create CUDA events ahead of time (scz_prof_reserve) so that a profiled, timed loop creates none"""
        self.check(self.L.scz_prof_reserve(self.h, C.c_uint64(int(events))))

    def prof_read(self, kernel_class):
        """-> (summed device ms, number of brackets) of one kernel class since prof_enable"""
        ms, n = C.c_double(), C.c_uint64()
        self.check(self.L.scz_prof_read(self.h, C.c_int32(self.KERNEL_CLASSES[kernel_class]), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def get_comm(self):
        """MPCNet::get_comm -> (upload, download) in the reference's serialised bytes"""
        up, down = C.c_uint64(), C.c_uint64()
        self.check(self.L.scz_ctx_get_comm(self.h, C.byref(up), C.byref(down)))
        return up.value, down.value

    # ---- device memory helpers (torch owns the allocations)
    def to_device(self, a, cols):
        a = _host(a, cols)
        return torch.from_numpy(a.view(np.int64)).to(self.device, non_blocking=False).contiguous()

    def empty(self, n, cols):
        return torch.empty((n, cols), dtype=torch.int64, device=self.device)

    @staticmethod
    def to_host(t):
        return t.detach().cpu().numpy().view(np.uint64)

    # ---- unit-level element-wise ops (device tensors)
    def fr_op(self, op, a, b):
        a, b = _dev(a, 4), _dev(b, 4)
        out = torch.empty_like(a)
        self.check(self.L.scz_fr_vec_op_dev(self.h, {"add": 0, "sub": 1, "mul": 2}[op], C.c_void_p(a.data_ptr()),
                                            C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), C.c_size_t(len(a))))
        return out

    def fq_op(self, op, a, b):
        a, b = _dev(a, 6), _dev(b, 6)
        out = torch.empty_like(a)
        self.check(self.L.scz_fq_vec_op_dev(self.h, {"add": 0, "sub": 1, "mul": 2}[op], C.c_void_p(a.data_ptr()),
                                            C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), C.c_size_t(len(a))))
        return out

    def _fr_unary(self, fn, a):
        a = _dev(a, 4)
        out = torch.empty_like(a)
        self.check(fn(self.h, C.c_void_p(a.data_ptr()), C.c_void_p(out.data_ptr()), C.c_size_t(len(a))))
        return out

    def fr_inv(self, a): return self._fr_unary(self.L.scz_fr_inv_dev, a)
    def fr_to_canonical(self, a): return self._fr_unary(self.L.scz_fr_to_canonical_dev, a)
    def fr_from_canonical(self, a): return self._fr_unary(self.L.scz_fr_from_canonical_dev, a)

    def g1_add_affine(self, acc_jac, aff, negate=None):
        acc, aff = _dev(acc_jac, 18), _dev(aff, 12)
        out = torch.empty_like(acc)
        neg = C.c_void_p(negate.data_ptr()) if negate is not None else None
        self.check(self.L.scz_g1_add_affine_dev(self.h, C.c_void_p(acc.data_ptr()), C.c_void_p(aff.data_ptr()), neg,
                                                C.c_void_p(out.data_ptr()), C.c_size_t(len(acc))))
        return out

    def g1_add(self, a, b):
        a, b = _dev(a, 18), _dev(b, 18)
        out = torch.empty_like(a)
        self.check(self.L.scz_g1_vec_op_dev(self.h, 0, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()),
                                            C.c_void_p(out.data_ptr()), C.c_size_t(len(a))))
        return out

    def g1_double(self, a):
        a = _dev(a, 18)
        out = torch.empty_like(a)
        self.check(self.L.scz_g1_vec_op_dev(self.h, 1, C.c_void_p(a.data_ptr()), None, C.c_void_p(out.data_ptr()),
                                            C.c_size_t(len(a))))
        return out

    def g1_mul(self, a, k):
        a, k = _dev(a, 18), _dev(k, 4)
        out = torch.empty_like(a)
        self.check(self.L.scz_g1_mul_fr_dev(self.h, C.c_void_p(a.data_ptr()), C.c_void_p(k.data_ptr()),
                                            C.c_void_p(out.data_ptr()), C.c_size_t(len(a))))
        return out

    def g1_to_affine(self, a):
        a = _dev(a, 18)
        out = self.empty(len(a), 12)
        self.check(self.L.scz_g1_to_affine_dev(self.h, C.c_void_p(a.data_ptr()), C.c_void_p(out.data_ptr()),
                                               C.c_size_t(len(a))))
        return out

    def g1_generator_mul(self, k):
        """synthetic bases k[i]*G (stands in for G1::rand, dpoly_comm.rs:214,229)"""
        k = _dev(k, 4)
        out = self.empty(len(k), 12)
        self.check(self.L.scz_g1_generator_mul_dev(self.h, C.c_void_p(k.data_ptr()), C.c_void_p(out.data_ptr()),
                                                   C.c_size_t(len(k))))
        return out

    # ---- wire formats (ark-serialize compressed)
    def g1_serialize_compressed(self, a):
        """(n, 18) Jacobian -> (n, 48) uint8: ark-bls12-381's compressed G1 encoding (Zcash format)"""
        a = _dev(a, 18)
        out = torch.empty((len(a), 48), dtype=torch.uint8, device=self.device)
        self.check(self.L.scz_g1_serialize_compressed_dev(self.h, C.c_void_p(a.data_ptr()), C.c_void_p(out.data_ptr()),
                                                          C.c_size_t(len(a))))
        return out

    def g1_deserialize_compressed(self, b):
        """(n, 48) uint8 -> ((n, 18) Jacobian, (n,) status: 0 ok, 1 malformed / off curve, 2 wrong subgroup)"""
        assert b.is_cuda and b.dtype == torch.uint8 and b.is_contiguous()
        b = b.reshape(-1, 48)
        out, st = self.empty(len(b), 18), torch.empty(len(b), dtype=torch.uint8, device=self.device)
        self.check(self.L.scz_g1_deserialize_compressed_dev(self.h, C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()),
                                                            C.c_void_p(st.data_ptr()), C.c_size_t(len(b))))
        return out, st

    def fr_deserialize(self, b):
        """(n, 4) int64 holding 32-byte little-endian canonical integers -> (Montgomery (n, 4), status)"""
        b = _dev(b, 4)
        out, st = torch.empty_like(b), torch.empty(len(b), dtype=torch.uint8, device=self.device)
        self.check(self.L.scz_fr_deserialize_dev(self.h, C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()),
                                                 C.c_void_p(st.data_ptr()), C.c_size_t(len(b))))
        return out, st

    def msm_set_window(self, c):
        self.check(self.L.scz_msm_set_window(self.h, C.c_uint32(c)))

    def msm_set_affine(self, mode=0, levels=0, slab_entries=0):
        """bucket accumulation variant (scz_msm_set_affine): mode 0 automatic, 1 always batched-affine, 2 never"""
        self.check(self.L.scz_msm_set_affine(self.h, C.c_uint32(mode), C.c_uint32(levels), C.c_uint64(slab_entries)))

    def msm_affine_sequences(self):
        n = C.c_uint64()
        self.check(self.L.scz_msm_affine_sequences(self.h, C.byref(n)))
        return n.value

    def msm_last_stats(self):
        a, b, w = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.check(self.L.scz_msm_last_stats(self.h, C.byref(a), C.byref(b), C.byref(w)))
        return {"bucket_adds": a.value, "buckets": b.value, "windows": w.value}


    def msm_use_precompute(self, on=True):
        """A/B switch: ignore the SRS fixed-base tables on this ctx when off"""
        self.check(self.L.scz_msm_use_precompute(self.h, C.c_int32(1 if on else 0)))

    def msm_cum_stats(self):
        a, p, q, g = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.check(self.L.scz_msm_cum_stats(self.h, C.byref(a), C.byref(p), C.byref(q), C.byref(g)))
        return {"bucket_adds": a.value, "pairs": p.value, "sequences": q.value, "segments": g.value}


# ---------------------------------------------------------------------------- G2 (d_msm is generic over CurveGroup, dmsm.rs:9-15)
G2_AFF_LIMBS, G2_JAC_LIMBS = 24, 36
# the BLS12-381 G2 generator, affine x | y with Fq2 = c0 | c1, Montgomery limbs (public constant)
G2_GENERATOR_AFFINE = np.array([[
    0xf5f28fa202940a10, 0xb3f5fb2687b4961a, 0xa1a893b53e2ae580, 0x9894999d1a3caee9, 0x6f67b7631863366b, 0x058191924350bcd7,
    0xa5a9c0759e23f606, 0xaaa0c59dbccd60c3, 0x3bb17e18e2867806, 0x1b1ab6cc8541b367, 0xc2b6ed0ef2158547, 0x11922a097360edf3,
    0x4c730af860494c4a, 0x597cfa1f5e369c5a, 0xe7e6856caa0a635a, 0xbbefb5e96e0d495f, 0x07d3a975f0ef25a2, 0x0083fd8e7e80dae5,
    0xadc0fc92df64b05d, 0x18aa270a2b1461dc, 0x86adac6a3be4eba0, 0x79495c4ec93da33a, 0xe7175850a43ccaed, 0x0b2bc2a163de1bf2]],
    dtype=np.uint64)
_FQ_ONE = np.array([0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745, 0x5c071a97a256ec6d,
                    0x15f65ec3fa80e493], dtype=np.uint64)


def g2_affine_to_jac(aff):
    """(n, 24) affine -> (n, 36) Jacobian with Z = 1 (identity rows, all zero, become (1, 1, 0))"""
    a = _host(aff, G2_AFF_LIMBS)
    out = np.zeros((len(a), G2_JAC_LIMBS), dtype=np.uint64)
    out[:, :24] = a
    out[:, 24:30] = _FQ_ONE
    inf = ~a.any(axis=1)
    out[inf, 0:6] = _FQ_ONE
    out[inf, 12:18] = _FQ_ONE
    out[inf, 24:30] = 0
    return out


def g2_op(ctx, op, a, b=None):
    """element-wise on Jacobian G2 points (device tensors (n, 36)): 'add' a + b, 'double' 2 a, 'mul' b * a with b Fr (n, 4)"""
    a = _dev(a, G2_JAC_LIMBS)
    out = torch.empty_like(a)
    code = {"add": 0, "double": 1, "mul": 2}[op]
    bp = None if b is None else _vp(_dev(b, G2_JAC_LIMBS if code == 0 else 4))
    ctx.check(ctx.L.scz_g2_vec_op_dev(ctx.h, C.c_int32(code), _vp(a), bp, _vp(out), C.c_size_t(len(a))))
    return out


def msm_g2(ctx, bases, scalars, inf_mask=None):
    """G2::msm(bases, scalars).  numpy: host path (scz_msm_g2; bases (n, 24) affine); torch: device path.  -> (1, 36) Jacobian"""
    if _is_dev(bases):
        return msm_g2_batched(ctx, [bases], [scalars])
    b, s = _host(bases, G2_AFF_LIMBS), _host(scalars, 4)
    out = np.zeros((1, G2_JAC_LIMBS), dtype=np.uint64)
    mask = None if inf_mask is None else np.ascontiguousarray(inf_mask, dtype=np.uint8)
    ctx.check(ctx.L.scz_msm_g2(ctx.h, b.ctypes.data_as(C.c_void_p), None if mask is None else mask.ctypes.data_as(C.c_void_p),
                               C.c_size_t(len(b)), s.ctypes.data_as(C.c_void_p), C.c_size_t(len(s)), out.ctypes.data_as(C.c_void_p)))
    return out


def msm_g2_batched(ctx, bases_list, scalars_list):
    k = len(bases_list)
    bl = [_dev(b, G2_AFF_LIMBS) for b in bases_list]
    sl = [_dev(s_, 4) for s_ in scalars_list]
    for b, s_ in zip(bl, sl):
        if len(b) != len(s_):
            raise SczError(-2, f"msm_g2: {len(b)} bases vs {len(s_)} scalars")
    out = ctx.empty(max(k, 1), G2_JAC_LIMBS)
    lens = (C.c_size_t * k)(*[len(b) for b in bl])
    ctx.check(ctx.L.scz_msm_g2_batched_dev(ctx.h, _ptr_array([b.data_ptr() for b in bl]), _ptr_array([s_.data_ptr() for s_ in sl]),
                                           lens, C.c_size_t(k), _vp(out)))
    return out[:k]


def d_msm_g2(ctx, pp, bases, scalars):
    """d_msm over G2 (dmsm.rs:9-43): device tensors, bases[k] (m_k, 24) affine, scalars[k] (m_k, 4) -> (batch, 36) Jacobian"""
    assert len(bases) == len(scalars)                                       # dmsm.rs:16
    k = len(bases)
    bl = [_dev(b, G2_AFF_LIMBS) for b in bases]
    sl = [_dev(s_, 4) for s_ in scalars]
    for b, s_ in zip(bl, sl):
        if len(b) != len(s_):
            raise SczError(-2, f"d_msm_g2: {len(b)} bases vs {len(s_)} scalars")   # G::msm(..).unwrap(), dmsm.rs:23
    out = ctx.empty(max(k, 1), G2_JAC_LIMBS)
    lens = (C.c_size_t * k)(*[len(b) for b in bl])
    ctx.check(ctx.L.scz_d_msm_g2_dev(ctx.h, pp.h, _ptr_array([b.data_ptr() for b in bl]), _ptr_array([s_.data_ptr() for s_ in sl]),
                                     lens, C.c_size_t(k), _vp(out)))
    return out[:k]


def d_msm_g2_leader(ctx, pp, gathered):
    """the leader closure (dmsm.rs:31-38) over G2 on a gathered buffer (n_parties, batch, 36) -> same shape"""
    g = ctx.to_device(np.ascontiguousarray(gathered, dtype=np.uint64).reshape(-1, G2_JAC_LIMBS), G2_JAC_LIMBS) if not _is_dev(gathered) else gathered
    n, batch = gathered.shape[0], gathered.shape[1]
    out = torch.empty_like(g)
    ctx.check(ctx.L.scz_d_msm_g2_leader_dev(ctx.h, pp.h, _vp(g), C.c_size_t(batch), _vp(out)))
    return ctx.to_host(out).reshape(n, batch, G2_JAC_LIMBS)


# ---------------------------------------------------------------------------- MSM
def msm(ctx, bases, scalars, inf_mask=None):
    """G1::msm(bases, scalars) (ark-ec VariableBaseMSM; dmsm.rs:23).  Raises SczError
    (SCZ_ERR_LEN_MISMATCH) where the reference's `.unwrap()` would panic."""
    if _is_dev(bases):
        b, s = _dev(bases, 12), _dev(scalars, 4)
        if len(b) != len(s):
            raise SczError(-2, f"msm: {len(b)} bases vs {len(s)} scalars")
        return msm_batched(ctx, [b], [s])
    b, s = _host(bases, 12), _host(scalars, 4)
    out = np.zeros((1, 18), dtype=np.uint64)
    mask = None
    if inf_mask is not None:
        mask = np.ascontiguousarray(inf_mask, dtype=np.uint8)
    ctx.check(ctx.L.scz_msm_g1(ctx.h, C.c_void_p(b.ctypes.data), C.c_void_p(mask.ctypes.data) if mask is not None else None,
                               C.c_size_t(len(b)), C.c_void_p(s.ctypes.data), C.c_size_t(len(s)),
                               C.c_void_p(out.ctypes.data)))
    return out


def msm_batched(ctx, bases_list, scalars_list):
    """device path: one launch sequence for a batch of MSMs -> (batch, 18) Jacobian tensor"""
    bs = [_dev(b, 12) for b in bases_list]
    ss = [_dev(s, 4) for s in scalars_list]
    for b, s in zip(bs, ss):
        if len(b) != len(s):
            raise SczError(-2, f"msm: {len(b)} bases vs {len(s)} scalars")
    k = len(bs)
    out = ctx.empty(k, 18)
    lens = (C.c_size_t * k)(*[len(s) for s in ss])
    ctx.check(ctx.L.scz_msm_g1_batched_dev(ctx.h, _ptr_array([b.data_ptr() for b in bs]),
                                           _ptr_array([s.data_ptr() for s in ss]), lens, C.c_size_t(k),
                                           C.c_void_p(out.data_ptr())))
    return out


# ---------------------------------------------------------------------------- PSS
class PackedSharingParams:
    """secret-sharing/src/pss.rs:17-171.  kind: 'fr' or 'g1' (Jacobian), like the
    reference's `G: DomainCoeff<F>`.  Operands are (batch, len, limbs) or (len, limbs)."""

    def __init__(self, ctx, l):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx.L.scz_pp_new(ctx.h, C.c_size_t(l), C.byref(h)))
        self.h = h
        self.l, self.n, self.t = l, 8 * l, l - 1

    def __del__(self):
        try:
            if self.h:
                self.ctx.L.scz_pp_free(self.h)
                self.h = None
        except Exception:
            pass

    def _apply(self, which, x, kind, len_in, len_out):
        ctx = self.ctx
        cols = 4 if kind == "fr" else 18
        host = not _is_dev(x)
        xd = ctx.to_device(x, cols) if host else _dev(x, cols)
        assert len(xd) % len_in == 0, f"operand length {len(xd)} is not a multiple of {len_in}"
        batch = len(xd) // len_in
        out = ctx.empty(batch * len_out, cols)
        k = 0 if kind == "fr" else 1
        a = (ctx.h, self.h, C.c_int32(k), C.c_void_p(xd.data_ptr()))
        if which == "pack":
            rc = ctx.L.scz_pss_pack_from_public_dev(*a, C.c_size_t(len_in), C.c_size_t(batch), C.c_void_p(out.data_ptr()))
        elif which == "pack_single":
            rc = ctx.L.scz_pss_pack_single_dev(*a, C.c_size_t(batch), C.c_void_p(out.data_ptr()))
        elif which == "unpack":
            rc = ctx.L.scz_pss_unpack_dev(*a, C.c_size_t(batch), C.c_void_p(out.data_ptr()))
        else:
            rc = ctx.L.scz_pss_unpack2_dev(*a, C.c_size_t(batch), C.c_void_p(out.data_ptr()))
        ctx.check(rc)
        out = out.reshape(batch, len_out, cols)
        return ctx.to_host(out) if host else out

    def pack_from_public(self, secrets, kind="fr", len_in=None):
        return self._apply("pack", secrets, kind, len_in or self.l, self.n)

    def pack_single(self, secret, kind="fr"):
        return self._apply("pack_single", secret, kind, 1, self.n)

    def unpack(self, shares, kind="fr"):
        return self._apply("unpack", shares, kind, self.n, self.l)

    def unpack2(self, shares, kind="fr"):
        return self._apply("unpack2", shares, kind, self.n, self.l)


# ---------------------------------------------------------------------------- d_msm
def d_msm(ctx, pp, bases, scalars):
    """dist-primitive/src/dmsm.rs:9-43: bases / scalars are lists (the batch) of arrays.
    Returns this party's packed shares of the batch results, (batch, 18) Jacobian."""
    if len(bases) != len(scalars):
        raise AssertionError("assert_eq!(bases.len(), scalars.len())")   # dmsm.rs:16
    k = len(bases)
    if k and _is_dev(bases[0]):
        bs = [_dev(b, 12) for b in bases]
        ss = [_dev(s, 4) for s in scalars]
        for b, s in zip(bs, ss):
            if len(b) != len(s):
                raise SczError(-2, f"d_msm: {len(b)} bases vs {len(s)} scalars")
        out = ctx.empty(k, 18)
        lens = (C.c_size_t * k)(*[len(s) for s in ss])
        ctx.check(ctx.L.scz_d_msm_dev(ctx.h, pp.h, _ptr_array([b.data_ptr() for b in bs]),
                                      _ptr_array([s.data_ptr() for s in ss]), lens, C.c_size_t(k),
                                      C.c_void_p(out.data_ptr())))
        return out
    bs = [_host(b, 12) for b in bases]
    ss = [_host(s, 4) for s in scalars]
    out = np.zeros((k, 18), dtype=np.uint64)
    bl = (C.c_size_t * k)(*[len(b) for b in bs])
    sl = (C.c_size_t * k)(*[len(s) for s in ss])
    ctx.check(ctx.L.scz_d_msm(ctx.h, pp.h, _ptr_array([b.ctypes.data for b in bs]), bl,
                              _ptr_array([s.ctypes.data for s in ss]), sl, C.c_size_t(k), C.c_void_p(out.ctypes.data)))
    return out


def d_msm_leader(ctx, pp, gathered):
    """the leader closure of d_msm alone (dmsm.rs:31-38): gathered is (n_parties, batch, 18) Jacobian,
    party-major; returns the (n_parties, batch, 18) buffer the leader scatters"""
    g = gathered if _is_dev(gathered) else ctx.to_device(np.ascontiguousarray(gathered).reshape(-1, 18), 18)
    g = g.reshape(pp.n, -1, 18).contiguous()
    batch = g.shape[1]
    out = torch.empty_like(g)
    ctx.check(ctx.L.scz_d_msm_leader_dev(ctx.h, pp.h, C.c_void_p(g.data_ptr()), C.c_size_t(batch),
                                         C.c_void_p(out.data_ptr())))
    return out if _is_dev(gathered) else ctx.to_host(out.reshape(-1, 18)).reshape(pp.n, batch, 18)


# ---------------------------------------------------------------------------- Fr tables
def _vp(t):
    return C.c_void_p(t.data_ptr())


def _in(ctx, x, cols):
    """-> (device tensor, was_host)"""
    if _is_dev(x):
        return _dev(x, cols), False
    return ctx.to_device(x, cols), True


def _out(ctx, t, host):
    return ctx.to_host(t) if host else t


def _log2(v):
    return v.bit_length() - 1


def fr_pointwise(ctx, mode, a, b, k=None):
    """dhyperplonk.rs:233-238, 251-256, 326-339.  mode: 'add' a+b, 'rsub' b-a, 'axpb' a + k[0]*b + k[1], 'div' a/b"""
    ad, host = _in(ctx, a, 4)
    bd, _ = _in(ctx, b, 4)
    kd = _in(ctx, k, 4)[0] if k is not None else None
    out = torch.empty_like(ad)
    m = {"add": 0, "rsub": 1, "axpb": 2, "div": 3}[mode]
    ctx.check(ctx.L.scz_fr_pointwise_dev(ctx.h, C.c_int32(m), _vp(ad), _vp(bd), _vp(kd) if kd is not None else None,
                                         _vp(out), C.c_size_t(len(ad))))
    if host and m == 3:   # host path: synchronous anyway, so the reference's panic on a zero denominator surfaces here
        ctx.raise_on_status()
    return _out(ctx, out, host)


def fix_variable(ctx, evals, points):
    """mle.rs:88-104"""
    ed, host = _in(ctx, evals, 4)
    pd, _ = _in(ctx, points, 4)
    k = min(len(pd), _log2(len(ed)))
    out = ctx.empty(len(ed) >> k, 4)
    ctx.check(ctx.L.scz_fix_variable_dev(ctx.h, _vp(ed), C.c_size_t(len(ed)), _vp(pd), C.c_size_t(len(pd)), _vp(out)))
    return _out(ctx, out, host)


def acc_product_tree(ctx, x):
    """the 2m-entry table of acc_product (dacc_product.rs:30-39)"""
    xd, host = _in(ctx, x, 4)
    out = ctx.empty(2 * len(xd), 4)
    ctx.check(ctx.L.scz_acc_product_dev(ctx.h, _vp(xd), C.c_size_t(len(xd)), _vp(out)))
    return _out(ctx, out, host)


def d_acc_product(ctx, x):
    """dacc_product.rs:365-414 -> (subtree, leader_tree or None)"""
    xd, host = _in(ctx, x, 4)
    sub = ctx.empty(2 * len(xd), 4)
    lead = ctx.empty(2 * ctx.n_parties, 4) if ctx.party_id == 0 else None
    ctx.check(ctx.L.scz_d_acc_product_dev(ctx.h, _vp(xd), C.c_size_t(len(xd)), _vp(sub),
                                          _vp(lead) if lead is not None else None))
    return _out(ctx, sub, host), (_out(ctx, lead, host) if lead is not None else None)


def sumcheck_rounds(ctx, f, g, challenge):
    """the local round loop alone -> (n triples, (f_last, g_last))"""
    fd, host = _in(ctx, f, 4)
    gd, _ = _in(ctx, g, 4)
    cd, _ = _in(ctx, challenge, 4)
    n = _log2(len(fd))
    out = ctx.empty(max(n, 1) * 3, 4)
    last = ctx.empty(2, 4)
    ctx.check(ctx.L.scz_sumcheck_product_rounds_dev(ctx.h, _vp(fd), _vp(gd), C.c_size_t(len(fd)), _vp(cd), _vp(out), _vp(last)))
    return _out(ctx, out[: n * 3].reshape(n, 3, 4), host), _out(ctx, last, host)


def sumcheck_product(ctx, f, g, challenge):
    """dsumcheck.rs:28-90 -> (n + 1, 3, 4)"""
    fd, host = _in(ctx, f, 4)
    gd, _ = _in(ctx, g, 4)
    cd, _ = _in(ctx, challenge, 4)
    n = _log2(len(fd))
    out = ctx.empty((n + 1) * 3, 4)
    ctx.check(ctx.L.scz_sumcheck_product_dev(ctx.h, _vp(fd), _vp(gd), C.c_size_t(len(fd)), _vp(cd), _vp(out)))
    return _out(ctx, out.reshape(n + 1, 3, 4), host)


def c_sumcheck_product(ctx, pp, f, g, challenge):
    """dsumcheck.rs:148-285 -> (n + log2 l + 1, 3, 4)"""
    fd, host = _in(ctx, f, 4)
    gd, _ = _in(ctx, g, 4)
    cd, _ = _in(ctx, challenge, 4)
    cnt = _log2(len(fd)) + _log2(pp.l) + 1
    out = ctx.empty(cnt * 3, 4)
    ctx.check(ctx.L.scz_c_sumcheck_product_dev(ctx.h, pp.h, _vp(fd), _vp(gd), C.c_size_t(len(fd)), _vp(cd), _vp(out)))
    return _out(ctx, out.reshape(cnt, 3, 4), host)


def d_sumcheck_product(ctx, f, g, challenge):
    """dsumcheck.rs:359-512 -> leader: (n + log2 N, 3, 4); every other party: an empty array (:507-509)"""
    fd, host = _in(ctx, f, 4)
    gd, _ = _in(ctx, g, 4)
    cd, _ = _in(ctx, challenge, 4)
    cap = _log2(len(fd)) + _log2(ctx.n_parties)
    out = ctx.empty(max(cap, 1) * 3, 4)
    cnt = C.c_size_t()
    ctx.check(ctx.L.scz_d_sumcheck_product_dev(ctx.h, _vp(fd), _vp(gd), C.c_size_t(len(fd)), _vp(cd), _vp(out), C.byref(cnt)))
    return _out(ctx, out[: cnt.value * 3].reshape(cnt.value, 3, 4), host)


def sumcheck(ctx, f, challenge):
    """dsumcheck.rs:6-26 -> (n + 1, 2, 4)"""
    fd, host = _in(ctx, f, 4)
    cd, _ = _in(ctx, challenge, 4)
    n = _log2(len(fd))
    out = ctx.empty((n + 1) * 2, 4)
    ctx.check(ctx.L.scz_sumcheck_dev(ctx.h, _vp(fd), C.c_size_t(len(fd)), _vp(cd), _vp(out)))
    return _out(ctx, out.reshape(n + 1, 2, 4), host)


def c_sumcheck(ctx, pp, f, challenge):
    """dsumcheck.rs:92-146 -> (n + log2 l + 1, 2, 4)"""
    fd, host = _in(ctx, f, 4)
    cd, _ = _in(ctx, challenge, 4)
    cnt = _log2(len(fd)) + _log2(pp.l) + 1
    out = ctx.empty(cnt * 2, 4)
    ctx.check(ctx.L.scz_c_sumcheck_dev(ctx.h, pp.h, _vp(fd), C.c_size_t(len(fd)), _vp(cd), _vp(out)))
    return _out(ctx, out.reshape(cnt, 2, 4), host)


def d_sumcheck(ctx, f, challenge):
    """dsumcheck.rs:287-357 -> leader: (n + log2 N, 2, 4); every other party: an empty array (:351-353)"""
    fd, host = _in(ctx, f, 4)
    cd, _ = _in(ctx, challenge, 4)
    cap = _log2(len(fd)) + _log2(ctx.n_parties)
    out = ctx.empty(max(cap, 1) * 2, 4)
    cnt = C.c_size_t()
    ctx.check(ctx.L.scz_d_sumcheck_dev(ctx.h, _vp(fd), C.c_size_t(len(fd)), _vp(cd), _vp(out), C.byref(cnt)))
    return _out(ctx, out[: cnt.value * 2].reshape(cnt.value, 2, 4), host)


def pss2ss(ctx, pp, share):
    """unpack.rs:72-97: one share -> Vec<F> of length l"""
    sd, host = _in(ctx, share, 4)
    out = ctx.empty(pp.l, 4)
    ctx.check(ctx.L.scz_pss2ss_dev(ctx.h, pp.h, _vp(sd), _vp(out)))
    return _out(ctx, out, host)


def degree_reduce(ctx, pp, share):
    """degree_reduce.rs:29-41"""
    sd, host = _in(ctx, share, 4)
    out = ctx.empty(1, 4)
    ctx.check(ctx.L.scz_degree_reduce_dev(ctx.h, pp.h, _vp(sd), _vp(out)))
    return _out(ctx, out, host)


class PolynomialCommitment:
    """The G1 side of dpoly_comm.rs:30-34 (`powers_of_g`) with the commit / open family (:236-464).
    levels[i]: packed affine bases of level i, numpy (n_i, 12) or CUDA tensors."""

    def __init__(self, ctx, levels):
        self.ctx = ctx
        self.levels = [lv if _is_dev(lv) else ctx.to_device(lv, 12) for lv in levels]   # keeps the arrays alive
        k = len(self.levels)
        ptrs = (C.c_void_p * k)(*[lv.data_ptr() for lv in self.levels])
        lens = (C.c_size_t * k)(*[len(lv) for lv in self.levels])
        h = C.c_void_p()
        ctx.check(ctx.L.scz_srs_from_device_levels(ctx.h, C.c_size_t(k), ptrs, lens, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h:
                self.ctx.L.scz_srs_free(self.h)
                self.h = None
        except Exception:
            pass

    @classmethod
    def _wrap(cls, ctx, handle):
        self = cls.__new__(cls)
        self.ctx, self.levels, self.h = ctx, [], handle
        return self

    @classmethod
    def new(cls, ctx, g_jac, s):
        """PolynomialCommitmentCub::new(g, _, s).mature() (dpoly_comm.rs:37-67): the real SRS for the trapdoor s"""
        g, _ = _in(ctx, g_jac, 18)
        sd, _ = _in(ctx, s, 4)
        h = C.c_void_p()
        ctx.check(ctx.L.scz_srs_new_dev(ctx.h, _vp(g), _vp(sd), C.c_size_t(len(sd)), C.byref(h)))
        return cls._wrap(ctx, h)

    def to_packed(self, pp, party):
        """to_packed (dpoly_comm.rs:164-194): party `party`'s PSS share of this SRS"""
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.scz_srs_to_packed_dev(self.ctx.h, self.h, pp.h, C.c_uint32(party), C.byref(h)))
        return PolynomialCommitment._wrap(self.ctx, h)

    def level(self, i):
        """packed affine points of level i as a CUDA tensor (len, 12)"""
        n = C.c_size_t()
        self.ctx.check(self.ctx.L.scz_srs_level_dev(self.ctx.h, self.h, C.c_size_t(i), None, C.byref(n)))
        out = self.ctx.empty(n.value, 12)
        self.ctx.check(self.ctx.L.scz_srs_level_dev(self.ctx.h, self.h, C.c_size_t(i), _vp(out), C.byref(n)))
        return out

    def precompute(self):
        """build the fixed-base tables of every level (scz_srs_precompute): same results, fewer bucket additions"""
        self.ctx.check(self.ctx.L.scz_srs_precompute(self.ctx.h, self.h))
        return self

    def commit(self, peval):
        ctx = self.ctx
        pd, host = _in(ctx, peval, 4)
        out = ctx.empty(1, 18)
        ctx.check(ctx.L.scz_commit_dev(ctx.h, self.h, _vp(pd), C.c_size_t(len(pd)), _vp(out)))
        return _out(ctx, out, host)

    def c_commit(self, pp, pevals):
        ctx = self.ctx
        ins = [_in(ctx, p, 4) for p in pevals]
        host = bool(ins) and ins[0][1]
        k = len(ins)
        out = ctx.empty(k, 18)
        ptrs = (C.c_void_p * k)(*[t.data_ptr() for t, _ in ins])
        lens = (C.c_size_t * k)(*[len(t) for t, _ in ins])
        ctx.check(ctx.L.scz_c_commit_dev(ctx.h, self.h, pp.h, ptrs, lens, C.c_size_t(k), _vp(out)))
        return _out(ctx, out, host)

    def d_commit(self, peval):
        ctx = self.ctx
        pd, host = _in(ctx, peval, 4)
        out = ctx.empty(1, 18)
        ctx.check(ctx.L.scz_d_commit_dev(ctx.h, self.h, _vp(pd), C.c_size_t(len(pd)), _vp(out)))
        return _out(ctx, out, host)

    def open(self, peval, point):
        ctx = self.ctx
        pd, host = _in(ctx, peval, 4)
        ud, _ = _in(ctx, point, 4)
        n = _log2(len(pd))
        val, proofs = ctx.empty(1, 4), ctx.empty(max(n, 1), 18)
        ctx.check(ctx.L.scz_open_dev(ctx.h, self.h, _vp(pd), C.c_size_t(len(pd)), _vp(ud), _vp(val), _vp(proofs)))
        return _out(ctx, val, host), _out(ctx, proofs[:n], host)

    def c_open(self, pp, peval, point):
        ctx = self.ctx
        pd, host = _in(ctx, peval, 4)
        ud, _ = _in(ctx, point, 4)
        cnt = _log2(len(pd)) + _log2(pp.l)
        val, proofs = ctx.empty(1, 4), ctx.empty(max(cnt, 1), 18)
        ctx.check(ctx.L.scz_c_open_dev(ctx.h, self.h, pp.h, _vp(pd), C.c_size_t(len(pd)), _vp(ud), _vp(val), _vp(proofs)))
        return _out(ctx, val, host), _out(ctx, proofs[:cnt], host)

    def d_open(self, peval, point):
        ctx = self.ctx
        pd, host = _in(ctx, peval, 4)
        ud, _ = _in(ctx, point, 4)
        cap = _log2(len(pd)) + _log2(ctx.n_parties)
        val, proofs = ctx.empty(1, 4), ctx.empty(max(cap, 1), 18)
        cnt = C.c_size_t()
        ctx.check(ctx.L.scz_d_open_dev(ctx.h, self.h, _vp(pd), C.c_size_t(len(pd)), _vp(ud), C.c_size_t(len(ud)), _vp(val),
                                       _vp(proofs), C.byref(cnt)))
        return _out(ctx, val, host), _out(ctx, proofs[: cnt.value], host)


# ---------------------------------------------------------------------------- the prover
class HpPk(C.Structure):
    """scz_hp_pk (include/scz.h)"""
    _TABLES = ("V", "a_evals", "b_evals", "c_evals", "I", "S1", "S2", "I_p", "S1_p", "S2_p", "ssigma_p", "sid_p", "eq",
               "eq_r1_p", "eq_r2_p", "challenge", "challenge_r1", "challenge_r2", "alpha_beta")
    _INJECTED = ("local_s_p", "local_s", "eq_leader")
    _fields_ = ([(k, C.c_void_p) for k in _TABLES] + [("c_commitment", C.c_void_p), ("d_commitment", C.c_void_p)]
                + [(k, C.c_void_p) for k in _INJECTED])


class HpItem(C.Structure):
    """scz_hp_item (include/scz.h)"""
    _fields_ = [(k, C.c_uint32) for k in ("kind", "triples_off", "triples_cnt", "points_off", "points_cnt", "value_off",
                                           "value_cnt")]


def hp_table_sizes(n, l, n_parties):
    """lengths of the tables `dhyperplonk` reads: PackedProvingParameters::new (dhyperplonk.rs:65-157) and :188-190"""
    gc = 1 << n
    return {"V": gc * 4 // l, "a_evals": gc // l, "b_evals": gc // l, "c_evals": gc // l, "I": gc // l, "S1": gc // l,
            "S2": gc // l, "I_p": gc // n_parties, "S1_p": gc // n_parties, "S2_p": gc // n_parties,
            "ssigma_p": gc * 4 // n_parties, "sid_p": gc * 4 // n_parties, "eq": gc // l, "eq_r1_p": gc * 4 // n_parties,
            "eq_r2_p": gc * 4 // n_parties, "challenge": n, "challenge_r1": n + 2, "challenge_r2": n + 2, "alpha_beta": 2,
            "local_s_p": gc * 4 // n_parties, "local_s": gc * 4 // n_parties // l, "eq_leader": 8 * l}


class PackedProvingParameters:
    """The fields of PackedProvingParameters (hyperplonk/src/dhyperplonk.rs:22-62) that `dhyperplonk` reads, as
    device-resident Fr tables, plus the three vectors the reference draws from entropy inside the function
    (:188-190), here explicit inputs.  `tables`: dict name -> (len, 4) array (numpy or CUDA tensor), names and
    lengths as hp_table_sizes; c_commitment / d_commitment: PolynomialCommitment (:53-54)."""

    def __init__(self, ctx, n, l, tables, c_commitment, d_commitment, data_parallel=False):
        self.ctx, self.n, self.l = ctx, n, l
        want = hp_table_sizes(n, l, ctx.n_parties)
        if data_parallel:   # dhyperplonk_data_parallel draws the whole s locally (dhyperplonk.rs:603)
            want["local_s"] = (1 << n) * 4 // l
        self.t = {}
        for name, ln in want.items():
            x = tables[name]
            x = _dev(x, 4) if _is_dev(x) else ctx.to_device(x, 4)
            if len(x) != ln:
                raise SczError(-2, f"PackedProvingParameters: {name} has {len(x)} entries, {ln} expected")
            self.t[name] = x
        self.c_commitment, self.d_commitment = c_commitment, d_commitment

    @classmethod
    def new(cls, ctx, n, l, seed=0, shared_seed=None, precompute=False):
        """PackedProvingParameters::new(n, l, pp) (:65-157) with synthetic random tables generated ON the device
        (the reference draws them with F::rand from entropy) and random-point SRS levels (new_single / new_random,
        dpoly_comm.rs:196-233: random points, not a valid SRS).  Challenges, alpha and beta come from `shared_seed`
        so that all parties agree on them."""
        N = ctx.n_parties
        gen = torch.Generator(device=ctx.device).manual_seed(0x5CA1AB1E ^ (seed << 32))
        pub = torch.Generator(device=ctx.device).manual_seed(0x5CA1AB1E ^ ((shared_seed if shared_seed is not None else seed) << 32) ^ 0xFFFF)

        def rand_fr(m, g):
            # uniform 254-bit integers used directly as Montgomery limbs: every value is a valid element (< r)
            t = torch.randint(-2**63, 2**63 - 1, (m, 4), dtype=torch.int64, device=ctx.device, generator=g)
            t[:, 3] &= (1 << 62) - 1
            return t
        tables = {}
        for name, ln in hp_table_sizes(n, l, N).items():
            public = name in ("challenge", "challenge_r1", "challenge_r2", "alpha_beta")
            tables[name] = rand_fr(ln, pub if public else gen)
        # a, b, c are fix_variable(V, .) of the witness table (:71-73): bind the two top variables to (0,0), (0,1), (1,0)
        zero, one = torch.zeros((1, 4), dtype=torch.int64, device=ctx.device), _fr_one(ctx)
        for name, pt in (("a_evals", (zero, zero)), ("b_evals", (zero, one)), ("c_evals", (one, zero))):
            tables[name] = fix_variable(ctx, tables["V"], torch.cat(pt))

        def levels(sizes):
            return [ctx.g1_generator_mul(rand_fr(m, gen)) for m in sizes]
        csz = [max(1, (1 << i) // l) for i in range(n + 3)]                      # new_single(n + 2, pp)
        dsz = [1 << i for i in range(n + 2 - (N.bit_length() - 1) + 1)]          # new_random(n + 2, N)
        c_srs, d_srs = PolynomialCommitment(ctx, levels(csz)), PolynomialCommitment(ctx, levels(dsz))
        if precompute:   # fixed-base tables: part of the proving key, built once (csrc/srs.cu)
            c_srs.precompute()
            d_srs.precompute()
        return cls(ctx, n, l, tables, c_srs, d_srs)

    def upload(self, host_tables):
        """copy HOST tables (name -> pinned int64 tensor or numpy array) into the resident device tables, on the ctx
        stream: the per-proof host -> device traffic of a caller whose witness lives in host memory"""
        for name, src in host_tables.items():
            dst = self.t[name]
            if not isinstance(src, torch.Tensor):
                src = torch.from_numpy(np.ascontiguousarray(src, dtype=np.uint64).view(np.int64))
            dst.copy_(src.view(dst.shape), non_blocking=True)

    def c_struct(self):
        pk = HpPk()
        for name in HpPk._TABLES + HpPk._INJECTED:
            setattr(pk, name, self.t[name].data_ptr())
        pk.c_commitment, pk.d_commitment = self.c_commitment.h, self.d_commitment.h
        return pk


def _fr_one(ctx):
    one = np.array([[0x00000001fffffffe, 0x5884b7fa00034802, 0x998c4fefecbc4ff5, 0x1824b159acc5056f]], dtype=np.uint64)
    return ctx.to_device(one, 4)


class ProofReader:
    """Device -> host read-back of proofs without stalling the prover's stream: the copy of proof i runs on its own
    stream behind an event, into pinned staging buffers, together with a snapshot of the ctx's status bits
    (scz_ctx_status_snapshot_dev) taken in stream order right after the proof.  `depth` proofs may be outstanding."""

    def __init__(self, ctx, depth=2):
        self.ctx, self.depth = ctx, depth
        with torch.cuda.device(ctx.device):
            self.stream = torch.cuda.Stream()
        self.slots = [None] * depth
        self.next = 0

    def submit(self, proof):
        ctx = self.ctx
        t, p, v = proof.used()
        slot = self.next % self.depth
        self.next += 1
        views = (proof.triples[: 3 * t], proof.points[:p], proof.values[:v])
        st = self.slots[slot]
        if st is not None and st.get("keep") is not None:
            raise RuntimeError(f"ProofReader: {self.depth} proofs already outstanding -- collect() the oldest ticket first "
                               "(its staging buffers would be overwritten)")
        if st is None or any(h.shape != d.shape for h, d in zip(st["host"], views)):
            st = {"host": [torch.empty(d.shape, dtype=d.dtype).pin_memory() for d in views],
                  "bits_host": torch.zeros(1, dtype=torch.int32).pin_memory(),
                  "bits": torch.zeros(1, dtype=torch.int32, device=ctx.device),
                  "ready": torch.cuda.Event(), "done": torch.cuda.Event()}
            self.slots[slot] = st
        ctx.check(ctx.L.scz_ctx_status_snapshot_dev(ctx.h, C.c_void_p(st["bits"].data_ptr())))
        with torch.cuda.device(ctx.device):
            st["ready"].record(ctx.stream)           # the stream the proof was enqueued on, whatever the caller's current one
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(st["ready"])
                for h, d in zip(st["host"], views):
                    h.copy_(d, non_blocking=True)
                st["bits_host"].copy_(st["bits"], non_blocking=True)
                st["done"].record(self.stream)
        st["keep"] = proof                      # the arenas stay alive until the copy has run
        return slot

    def collect(self, slot):
        st = self.slots[slot]
        if st is None or st.get("keep") is None:
            raise RuntimeError("ProofReader.collect: no outstanding proof behind this ticket (collected twice?)")
        st["done"].synchronize()
        st["keep"] = None
        if int(st["bits_host"][0]) & 1:
            raise ZeroDivisionError("field division by zero (arkworks panics here: hyperplonk/src/dhyperplonk.rs:338-339)")
        return tuple(h.numpy().view(np.uint64).copy() for h in st["host"])


class HyperPlonkProof:
    """The return value of `dhyperplonk` (dhyperplonk.rs:567-570) on the device: three arenas + the item table.
    `nested()` rebuilds the reference's tuple
        ((gate_identity_proofs, gate_identity_commitments), (wiring_proofs, wiring_commits, wiring_opens))
    with numpy arrays: proofs (cnt, 3, 4); commitments (1, 18); opens (value (1, 4), proofs (k, 18))."""

    def __init__(self, ctx, triples, points, values, items):
        self.ctx, self.triples, self.points, self.values, self.items = ctx, triples, points, values, items

    def used(self):
        """(triples, points, values) element counts actually written"""
        t = max([it.triples_off + it.triples_cnt for it in self.items] + [0])
        p = max([it.points_off + it.points_cnt for it in self.items] + [0])
        v = max([it.value_off + it.value_cnt for it in self.items] + [0])
        return t, p, v

    def to_host(self):
        """device -> host copy of the written part of the three arenas (numpy, synchronous); raises ZeroDivisionError
        when the proof divided by zero (the reference panics, dhyperplonk.rs:338-339)"""
        t, p, v = self.used()
        self.ctx.raise_on_status()
        return (self.ctx.to_host(self.triples[: 3 * t]), self.ctx.to_host(self.points[:p]), self.ctx.to_host(self.values[:v]))

    def to_host_async(self, reader):
        """the same copy queued on `reader`'s stream without waiting for it: `reader.collect(ticket)` returns the three
        arrays.  A host that keeps several proofs in flight reads proof i back while proof i + 1 is already enqueued
        (bench.py's e2e leg)."""
        return reader.submit(self)

    def nested(self):
        tri = self.ctx.to_host(self.triples).reshape(-1, 3, 4)
        pts = self.ctx.to_host(self.points)
        val = self.ctx.to_host(self.values)
        gp, gc, wp, wc, wo = [], [], [], [], []
        for it in self.items:
            t = tri[it.triples_off: it.triples_off + it.triples_cnt]
            p = pts[it.points_off: it.points_off + it.points_cnt]
            v = val[it.value_off: it.value_off + it.value_cnt]
            if it.kind == 0:
                gp.append(t)
            elif it.kind == 1:
                gc.append((p[:1], (v, p[1:])))
            elif it.kind == 2:
                wp.append(t)
            elif it.kind == 3:
                wc.append(p)
            else:
                wo.append((v, p))
        return (gp, gc), (wp, wc, wo)


def dhyperplonk(ctx, n, pk, pp, _entry="scz_dhyperplonk_dev"):
    """hyperplonk/src/dhyperplonk.rs:159-571 (after net.sync(), :193) -> HyperPlonkProof.  Asynchronous on the
    ctx stream in leader mode; call ctx.sync() or .nested() to wait."""
    L = ctx.L
    nt, npt, nv, ni = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    ctx.check(L.scz_dhyperplonk_sizes(C.c_size_t(n), C.c_size_t(pp.l), C.c_size_t(ctx.n_parties), C.byref(nt), C.byref(npt),
                                      C.byref(nv), C.byref(ni)))
    tri = ctx.empty(nt.value * 3, 4)
    pts = ctx.empty(npt.value, 18)
    val = ctx.empty(nv.value, 4)
    items = (HpItem * ni.value)()
    cnt = C.c_size_t()
    cpk = pk.c_struct()
    ctx.check(getattr(L, _entry)(ctx.h, C.c_size_t(n), C.byref(cpk), pp.h, _vp(tri), nt, _vp(pts), npt, _vp(val), nv,
                                 items, ni, C.byref(cnt)))
    return HyperPlonkProof(ctx, tri, pts, val, list(items[: cnt.value]))


class CpermPk(C.Structure):
    """scz_cperm_pk (include/scz.h)"""
    _TABLES = ("V", "sid", "ssigma", "eq_r1", "mask", "unmask0", "unmask1", "unmask2", "challenge_r1", "alpha_beta")
    _fields_ = [(k, C.c_void_p) for k in _TABLES] + [("c_commitment", C.c_void_p)]


def c_acc_product_and_share(ctx, pp, shares, masks, unmask0, unmask1, unmask2):
    """dacc_product.rs:66-292 -> (v(x,0), v(x,1), v(1,x)) shares, each as long as `shares`"""
    ins = [_in(ctx, x, 4) for x in (shares, masks, unmask0, unmask1, unmask2)]
    host = ins[0][1]
    n = len(ins[0][0])
    outs = [ctx.empty(n, 4) for _ in range(3)]
    ctx.check(ctx.L.scz_c_acc_product_and_share_dev(ctx.h, pp.h, *[_vp(t) for t, _ in ins], C.c_size_t(n),
                                                    *[_vp(t) for t in outs]))
    return tuple(_out(ctx, t, host) for t in outs)


def cpermcheck(ctx, n, tables, c_commitment, pp):
    """hyperplonk/src/dhyperplonk.rs:1249-1385.  tables: dict with the CpermPk._TABLES entries ((len, 4) arrays, numpy or
    CUDA tensors).  Returns a HyperPlonkProof whose nested() is ((), (wiring_proofs, wiring_commits, wiring_opens))."""
    L = ctx.L
    keep = {k: (_dev(tables[k], 4) if _is_dev(tables[k]) else ctx.to_device(tables[k], 4)) for k in CpermPk._TABLES}
    cpk = CpermPk()
    for k, t in keep.items():
        setattr(cpk, k, t.data_ptr())
    cpk.c_commitment = c_commitment.h
    nt, npt, nv, ni = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    ctx.check(L.scz_dhyperplonk_sizes(C.c_size_t(n), C.c_size_t(pp.l), C.c_size_t(ctx.n_parties), C.byref(nt), C.byref(npt),
                                      C.byref(nv), C.byref(ni)))
    tri, pts, val = ctx.empty(nt.value * 3, 4), ctx.empty(npt.value, 18), ctx.empty(nv.value, 4)
    items = (HpItem * ni.value)()
    cnt = C.c_size_t()
    ctx.check(L.scz_cpermcheck_dev(ctx.h, C.c_size_t(n), C.byref(cpk), pp.h, _vp(tri), nt, _vp(pts), npt, _vp(val), nv, items, ni,
                                   C.byref(cnt)))
    proof = HyperPlonkProof(ctx, tri, pts, val, list(items[: cnt.value]))
    proof._keep = keep
    return proof


class LocalPk(C.Structure):
    """scz_local_pk (include/scz.h)"""
    _TABLES = ("m", "a_evals", "b_evals", "c_evals", "input", "q1", "q2", "ssigma", "sid", "eq", "eq_p2", "challenge",
               "challengep2", "alpha_beta")
    _fields_ = [(k, C.c_void_p) for k in _TABLES] + [("commitment", C.c_void_p)]


def local_hyperplonk(ctx, n, tables, commitment):
    """hyperplonk/src/hyperplonk.rs:15-160 ("Local HyperPlonk", the monolithic baseline) on explicit inputs.
    tables: dict with the LocalPk._TABLES entries; commitment: PolynomialCommitment with levels 0 .. n+2."""
    keep = {k: (_dev(tables[k], 4) if _is_dev(tables[k]) else ctx.to_device(tables[k], 4)) for k in LocalPk._TABLES}
    cpk = LocalPk()
    for k, t in keep.items():
        setattr(cpk, k, t.data_ptr())
    cpk.commitment = commitment.h
    L = ctx.L
    nt, npt, nv, ni = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    ctx.check(L.scz_dhyperplonk_sizes(C.c_size_t(n), C.c_size_t(1), C.c_size_t(ctx.n_parties), C.byref(nt), C.byref(npt),
                                      C.byref(nv), C.byref(ni)))
    tri, pts, val = ctx.empty(nt.value * 3, 4), ctx.empty(npt.value, 18), ctx.empty(nv.value, 4)
    items = (HpItem * ni.value)()
    cnt = C.c_size_t()
    ctx.check(L.scz_local_hyperplonk_dev(ctx.h, C.c_size_t(n), C.byref(cpk), _vp(tri), nt, _vp(pts), npt, _vp(val), nv, items, ni,
                                         C.byref(cnt)))
    proof = HyperPlonkProof(ctx, tri, pts, val, list(items[: cnt.value]))
    proof._keep = keep
    return proof


def dhyperplonk_data_parallel(ctx, n, pk, pp):
    """dhyperplonk.rs:573-960: pk.t['local_s'] holds the whole `s` (4 * 2^n / l entries, :603); no step-2.a exchange"""
    return dhyperplonk(ctx, n, pk, pp, "scz_dhyperplonk_data_parallel_dev")


def dpermcheck(ctx, n, pk, pp):
    """dhyperplonk.rs:962-1247: the wiring identity alone; .nested() returns ((), (proofs, commits, opens)) with empty gate lists"""
    return dhyperplonk(ctx, n, pk, pp, "scz_dpermcheck_dev")


__all__ = ["Context", "PackedSharingParams", "msm", "msm_batched", "d_msm", "d_msm_leader", "NetVTable",
           "msm_g2", "msm_g2_batched", "d_msm_g2", "d_msm_g2_leader", "g2_op", "g2_affine_to_jac", "G2_GENERATOR_AFFINE",
           "fr_pointwise", "fix_variable", "acc_product_tree", "d_acc_product", "sumcheck_rounds", "sumcheck_product",
           "c_sumcheck_product", "d_sumcheck_product", "sumcheck", "c_sumcheck", "d_sumcheck", "pss2ss", "degree_reduce", "PolynomialCommitment",
           "PackedProvingParameters", "HyperPlonkProof", "dhyperplonk", "dhyperplonk_data_parallel", "dpermcheck", "cpermcheck", "c_acc_product_and_share", "local_hyperplonk",
           "hp_table_sizes"]
