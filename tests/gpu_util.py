"""Helpers shared by the -m gpu parity tests (CUDA path through the C ABI vs the oracle)."""
import numpy as np


def packed_affine(aff13):
    """oracle affine (n,13) -> product layout (n,12) with infinity folded into x = y = 0"""
    a = np.ascontiguousarray(aff13, dtype=np.uint64).reshape(-1, 13)
    out = a[:, :12].copy()
    out[(a[:, 12] & np.uint64(0xFFFFFFFF)) != 0] = 0
    return out


def oracle_affine(packed12):
    """product layout (n,12) -> oracle affine (n,13)"""
    p = np.ascontiguousarray(packed12, dtype=np.uint64).reshape(-1, 12)
    out = np.zeros((len(p), 13), dtype=np.uint64)
    out[:, :12] = p
    out[~p.any(axis=1), 12] = 1
    return out


def fr_sum(orc, a):
    """sum of an (n,4) Fr array via pairwise oracle additions"""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    while len(a) > 1:
        if len(a) % 2:
            a = np.concatenate([a, np.zeros((1, 4), dtype=np.uint64)])
        h = len(a) // 2
        a = orc.fr_add(a[:h], a[h:])
    return a


def fr_dot(orc, a, b):
    return fr_sum(orc, orc.fr_mul(a, b))
