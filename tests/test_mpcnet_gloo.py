"""CPU (no GPU): the byte-level MPCNet over torch.distributed/gloo with world_size 2 and 3, and the share-file
framing of the delegator (Vec<Fr>::serialize_uncompressed) against the big-int twin."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import py_twin as tw
from tests.test_net_gloo import ROOT, _free_port


@pytest.mark.parametrize("world", [2, 3])
def test_mpcnet_bytes_gloo(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_gloo_mpcnet_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GLOO_MPCNET_OK" in r.stdout


def test_share_file_framing(orc):
    from scz_b200.delegator import decode_vec_fr, encode_vec_fr
    vals = [0, 1, tw.R_MOD - 1, 0x0123456789abcdef_fedcba9876543210_0011223344556677 % tw.R_MOD, 1 << 200]
    canon = orc.ints_to_arr(vals, 4)
    data = encode_vec_fr(canon)
    # u64 LE length, then 32-byte little-endian canonical integers (ark-serialize 0.4.2, Vec<T> and Fp)
    assert data[:8] == (5).to_bytes(8, "little") and len(data) == 8 + 32 * 5
    assert [int.from_bytes(data[8 + 32 * i: 40 + 32 * i], "little") for i in range(5)] == vals
    assert np.array_equal(decode_vec_fr(data), canon)
    assert encode_vec_fr(np.zeros((0, 4), dtype=np.uint64)) == bytes(8)
    assert decode_vec_fr(bytes(8)).shape == (0, 4)
    for bad in (b"", data[:-1], data + b"\0", (6).to_bytes(8, "little") + data[8:]):
        with pytest.raises(ValueError):
            decode_vec_fr(bad)
