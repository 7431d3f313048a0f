"""-m gpu: the delegator's share files (dist-primitive/examples/delegator.rs) -- pack on the device, files in
ark-serialize's uncompressed Vec<Fr> format, read back, and the reference's own round-trip assertions."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import scz_b200 as scz
    c = scz.Context(device=0, n_parties=8)
    yield c
    c.close()


@pytest.mark.parametrize("l,n", [(1, 37), (2, 64), (2, 33), (4, 1000)])
def test_delegate_write_read(orc, ctx, tmp_path, l, n):
    import scz_b200 as scz
    from scz_b200.delegator import Delegator, decode_vec_fr, read_vec_fr
    rng = np.random.default_rng(900 + l)
    x = orc.random_fr(rng, n)
    pp = scz.PackedSharingParams(ctx, l)
    d = Delegator(ctx, x)
    workers = d.write(pp, str(tmp_path))
    chunks = (n + l - 1) // l
    assert tuple(workers.shape) == (8 * l, chunks, 4)
    # per-chunk oracle packing (pss.rs:69-73), zero-padded last chunk
    xp = np.concatenate([x, np.zeros((chunks * l - n, 4), dtype=np.uint64)])
    opp = orc.pp_new(l)
    want = np.stack([orc.pack_from_public(opp, xp[c * l:(c + 1) * l]).reshape(8 * l, 4) for c in range(chunks)], axis=1)
    assert np.array_equal(ctx.to_host(workers), want)
    # the files: delegator.rs:91-110's own assertions, plus the exact bytes
    names = sorted(os.listdir(tmp_path))
    assert names == sorted(["delegator"] + [f"worker_{j}" for j in range(8 * l)])
    raw = open(tmp_path / "delegator", "rb").read()
    assert [orc.limbs_to_int(r) for r in decode_vec_fr(raw)] == orc.fr_to_ints(x)
    assert np.array_equal(ctx.to_host(read_vec_fr(ctx, str(tmp_path / "delegator"))), x)
    for j in range(8 * l):
        assert np.array_equal(ctx.to_host(read_vec_fr(ctx, str(tmp_path / f"worker_{j}"))), want[j])
    # unpack of the shares gives the secrets back (pss.rs:202-214's property)
    back = pp.unpack(workers.permute(1, 0, 2).contiguous())
    assert np.array_equal(ctx.to_host(back).reshape(-1, 4), xp)


def test_delegator_errors(orc, ctx, tmp_path):
    import scz_b200 as scz
    from scz_b200.delegator import Delegator, read_vec_fr
    pp = scz.PackedSharingParams(ctx, 1)
    d = Delegator(ctx, orc.random_fr(np.random.default_rng(1), 4))
    with pytest.raises(FileNotFoundError):
        d.write(pp, str(tmp_path / "missing"))
    # a non-canonical element (>= r) is rejected on read, like ark-ff's deserialize
    bad = (1).to_bytes(8, "little") + b"\xff" * 32
    (tmp_path / "bad").write_bytes(bad)
    with pytest.raises(ValueError):
        read_vec_fr(ctx, str(tmp_path / "bad"))
