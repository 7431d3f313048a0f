"""The delegator's share files (dist-primitive/examples/delegator.rs:35-111; the same `Delegator::delegate` feeds
examples/sumcheck.rs:118,154,200,237): a witness vector x is cut into chunks of l secrets, every chunk is packed
with `pack_from_public` (pss.rs:69-73, on the device here, all chunks in one launch) and share j of every chunk
goes to worker j.  On disk: `<dir>/delegator` = x and `<dir>/worker_<j>` = worker j's shares, each a
`Vec<Fr>::serialize_uncompressed` (ark-serialize 0.4.2: u64 little-endian length, then 32-byte little-endian
canonical integers).  The output directory must exist (delegator.rs:77-79 panics otherwise).
"""
import os

import numpy as np


def encode_vec_fr(canon):
    """(n, 4) uint64 / int64 canonical little-endian limbs -> bytes of Vec<Fr>::serialize_uncompressed"""
    a = np.ascontiguousarray(canon).view(np.uint64).reshape(-1, 4)
    return len(a).to_bytes(8, "little") + a.astype("<u8").tobytes()


def decode_vec_fr(data):
    """inverse of encode_vec_fr (framing only: range validation is ctx.fr_deserialize's job)"""
    if len(data) < 8:
        raise ValueError("truncated Vec<Fr>: no length prefix")
    n = int.from_bytes(data[:8], "little")
    if len(data) != 8 + 32 * n:
        raise ValueError(f"Vec<Fr> of length {n} needs {8 + 32 * n} bytes, got {len(data)}")
    return np.frombuffer(data, dtype="<u8", offset=8).reshape(n, 4).copy()


class Delegator:
    """x: (n, 4) Montgomery Fr on the host or the device (delegator.rs:24-27)"""

    def __init__(self, ctx, x):
        self.ctx = ctx
        self.x = ctx.to_device(x, 4) if not hasattr(x, "is_cuda") else x.reshape(-1, 4)

    def delegate(self, pp):
        """-> (8l, ceil(n / l), 4) device tensor: row j = worker j's x_shares (delegator.rs:49-62).  A short last
        chunk is packed as it is in the reference: pack_from_public zero-pads to the secret domain (pss.rs:95)."""
        import torch
        n, l = len(self.x), pp.l
        chunks = (n + l - 1) // l
        x = self.x
        if chunks * l != n:
            x = torch.cat([x, torch.zeros((chunks * l - n, 4), dtype=x.dtype, device=x.device)])
        shares = pp.pack_from_public(x, kind="fr")          # (chunks, 8l, 4)
        return shares.permute(1, 0, 2).contiguous()

    def write(self, pp, out_dir):
        """writes `delegator` and `worker_<j>`; returns the per-worker share tensor"""
        if not os.path.isdir(out_dir):
            raise FileNotFoundError(f"{out_dir} does not exist")
        ctx = self.ctx
        workers = self.delegate(pp)
        with open(os.path.join(out_dir, "delegator"), "wb") as f:
            f.write(encode_vec_fr(ctx.to_host(ctx.fr_to_canonical(self.x))))
        canon = ctx.to_host(ctx.fr_to_canonical(workers.reshape(-1, 4))).reshape(workers.shape[0], -1, 4)
        for j in range(workers.shape[0]):
            with open(os.path.join(out_dir, f"worker_{j}"), "wb") as f:
                f.write(encode_vec_fr(canon[j]))
        return workers


def read_vec_fr(ctx, path):
    """`Vec::<Fr>::deserialize_uncompressed` of a share file -> (n, 4) Montgomery Fr on the device; a limb pattern
    >= r is an error, as in ark-ff"""
    with open(path, "rb") as f:
        canon = decode_vec_fr(f.read())
    if len(canon) == 0:
        return ctx.empty(0, 4)
    out, status = ctx.fr_deserialize(ctx.to_device(canon.view(np.int64), 4))
    if int(status.max().item()) != 0:
        raise ValueError(f"{path}: element {int(status.argmax().item())} is not a canonical Fr")
    return out
