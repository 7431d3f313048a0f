"""-m gpu: the collaborative HyperPlonk prover (`dhyperplonk`, hyperplonk/src/dhyperplonk.rs:159-571) through the
C ABI (scz_dhyperplonk_dev), entry by entry against the oracle's restatement (oracle/hyperplonk.py): leader mode on
one ctx and parties mode (N = 8 ctxs on one GPU under LocalTestNet)."""
import os

import numpy as np
import pytest

from tests.gpu_util import oracle_affine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _tables_for_product(pk):
    """oracle pk (alpha, beta separate) -> product tables (alpha_beta)"""
    t = {k: v for k, v in pk.items() if k not in ("c_commitment", "d_commitment", "alpha", "beta")}
    t["alpha_beta"] = np.concatenate([pk["alpha"], pk["beta"]])
    return t


def _make_srs(seed_ctx, orc, rng, sizes):
    dev, host = [], []
    for m in sizes:
        b = seed_ctx.g1_generator_mul(seed_ctx.to_device(orc.random_fr(rng, m), 4))
        dev.append(b)
        host.append(oracle_affine(seed_ctx.to_host(b)))
    return dev, orc.Srs.from_levels(host)


def _same_proof(orc, got, want, who):
    (gp, gc), (wp, wc, wo) = got
    assert len(gp) == len(want["gate_identity_proofs"]), who
    for a, b in zip(gp, want["gate_identity_proofs"]):
        assert np.array_equal(a, b), who
    assert len(gc) == len(want["gate_identity_commitments"]), who
    for (com, (val, proofs)), (ocom, (oval, oproofs)) in zip(gc, want["gate_identity_commitments"]):
        assert orc.canon_g1(com) == orc.canon_g1(ocom), who
        assert np.array_equal(val, oval), who
        assert len(proofs) == len(oproofs) and orc.canon_g1(proofs) == orc.canon_g1(oproofs), who
    assert len(wp) == len(want["wiring_proofs"]), (who, len(wp), len(want["wiring_proofs"]))
    for k, (a, b) in enumerate(zip(wp, want["wiring_proofs"])):
        assert a.shape == b.shape and np.array_equal(a, b), (who, k)
    assert len(wc) == len(want["wiring_commits"]), who
    for a, b in zip(wc, want["wiring_commits"]):
        assert orc.canon_g1(a) == orc.canon_g1(b), who
    assert len(wo) == len(want["wiring_opens"]), who
    for k, ((val, proofs), (oval, oproofs)) in enumerate(zip(wo, want["wiring_opens"])):
        assert np.array_equal(val, oval), (who, k)
        assert len(proofs) == len(oproofs) and orc.canon_g1(proofs) == orc.canon_g1(oproofs), (who, k)


# n >= 12: every sumcheck / fold starts in the multi-CTA kernels, the product tree reaches k_tree_level, the MSM batch
# mixes 2^0 .. 2^(n+2)-point segments with wide windows and 256-entry accumulate chunks -- the sizes at which the
# prover's integration (not only the per-kernel tests) is compared entry by entry with the oracle.
@pytest.mark.parametrize("n,l,pre", [(4, 1, False), (6, 1, False), (6, 2, False), (6, 1, True), (7, 2, True),
                                     (12, 1, False), (12, 2, True), (14, 1, True), (14, 2, True), (16, 1, True)])
def test_dhyperplonk_leader_mode(orc, n, l, pre):
    import scz_b200 as scz
    from oracle import hyperplonk as ohp
    N = 8 * l
    orc.set_msm_threads(os.cpu_count() or 1)   # the oracle's MSM windows over all host threads (same group elements)
    rng = np.random.default_rng(600 + 10 * n + l)
    ctx = scz.Context(device=0, n_parties=N)
    pp, opp = scz.PackedSharingParams(ctx, l), orc.pp_new(l)
    csz, dsz = ohp.srs_level_sizes(n, l, N)
    cdev, csrs = _make_srs(ctx, orc, rng, csz)
    ddev, dsrs = _make_srs(ctx, orc, rng, dsz)
    opk = ohp.random_pk(rng, n, l, N, csrs, dsrs)
    want = ohp.dhyperplonk(n, [opk], opp, orc.LEADER_SIM, N)[0]
    c_srs, d_srs = scz.PolynomialCommitment(ctx, cdev), scz.PolynomialCommitment(ctx, ddev)
    if pre:   # fixed-base tables (csrc/srs.cu): same group elements
        c_srs.precompute()
        d_srs.precompute()
    pk = scz.PackedProvingParameters(ctx, n, l, _tables_for_product(opk), c_srs, d_srs)
    got = scz.dhyperplonk(ctx, n, pk, pp).nested()
    _same_proof(orc, got, want, "leader")
    if pre:
        adds_pre = ctx.msm_cum_stats()["bucket_adds"]
        ctx.msm_use_precompute(False)
        _same_proof(orc, scz.dhyperplonk(ctx, n, pk, pp).nested(), want, "leader, tables ignored")
        assert ctx.msm_cum_stats()["bucket_adds"] - adds_pre != adds_pre    # the two runs really took different paths
    # shape of the reference's return value at l = 1 (dhyperplonk.rs:567-570)
    (gp, gc), (wp, wc, wo) = got
    s = N.bit_length() - 1
    assert len(gp) == 6 and len(gc) == 6
    assert len(wp) == 1 + 3 + 3 * (n - s) + 3 and len(wc) == 1 + 8 + 3 and len(wo) == 3 + 5 + 3 * (n - s) + 3
    up, down = ctx.get_comm()
    assert up > 0 and down > 0
    ctx.close()


def test_dhyperplonk_generated_parameters_round_identities(orc):
    """PackedProvingParameters.new (device-generated synthetic tables, a = fix_variable(V, (0,0)) etc.) at n = 10:
    every sumcheck proof of the gate identity satisfies the verifier's round identity (dsumcheck.rs:558-588) and
    the a / b / c tables are the right quarters of V (mle.rs:88-104 with points in {0, 1})."""
    import scz_b200 as scz
    from oracle import py_twin as tw
    n, l, N = 10, 1, 8
    ctx = scz.Context(device=0, n_parties=N)
    pp = scz.PackedSharingParams(ctx, l)
    pk = scz.PackedProvingParameters.new(ctx, n, l, seed=3)
    V = ctx.to_host(pk.t["V"])
    q = len(V) // 4
    assert np.array_equal(ctx.to_host(pk.t["a_evals"]), V[:q])          # (0, 0): top two variables zero
    assert np.array_equal(ctx.to_host(pk.t["b_evals"]), V[q:2 * q])     # (0, 1)
    assert np.array_equal(ctx.to_host(pk.t["c_evals"]), V[2 * q:3 * q])  # (1, 0)
    (gp, gc), (wp, wc, wo) = scz.dhyperplonk(ctx, n, pk, pp).nested()
    R = tw.R_MOD
    inv2 = pow(2, R - 2, R)
    chi = orc.fr_to_ints(ctx.to_host(pk.t["challenge"]))
    for proof in gp:
        tri = [[orc.fr_to_ints(proof[i, j:j + 1])[0] for j in range(3)] for i in range(len(proof))]
        for i in range(n - 1):
            p0, p1, p2 = tri[i]
            r = chi[i]
            val = (p0 * (r - 1) * (r - 2) * inv2 - p1 * r * (r - 2) + p2 * r * (r - 1) * inv2) % R
            assert val == (tri[i + 1][0] + tri[i + 1][1]) % R
    assert all(len(p) == n for _, (_, p) in gc)
    ctx.close()


@pytest.mark.parametrize("net_kind", ["local", "native_hub"])
def test_dhyperplonk_parties_mode(orc, net_kind):
    """N = 8 parties (l = 1) on one GPU; every party's proof against the oracle's 8-party run.  "local": LocalTestNet
    (Python callbacks); "native_hub": libscz's own hub (csrc/nccl_net.cu) with world = 1 and 8 parties per rank -- the
    host barrier, the event fences and the device copies of the native data plane, everything but NCCL itself."""
    import scz_b200 as scz
    from oracle import hyperplonk as ohp
    from scz_b200.net import LocalTestNet, NativeNcclNet
    n, l, N = 5, 1, 8
    rng = np.random.default_rng(640)
    opp = orc.pp_new(l)
    seed_ctx = scz.Context(device=0, n_parties=N)
    csz, dsz = ohp.srs_level_sizes(n, l, N)
    opks, srs_dev = [], []
    for j in range(N):
        cdev, csrs = _make_srs(seed_ctx, orc, rng, csz)
        ddev, dsrs = _make_srs(seed_ctx, orc, rng, dsz)
        opks.append(ohp.random_pk(rng, n, l, N, csrs, dsrs, shared=opks[0] if opks else None))
        srs_dev.append((cdev, ddev))
    want = ohp.dhyperplonk(n, opks, opp, orc.PARTIES, N)

    def party(j, net):
        c = scz.Context(device=0, party_id=j, n_parties=N, net=net)
        pp = scz.PackedSharingParams(c, l)
        pk = scz.PackedProvingParameters(c, n, l, _tables_for_product(opks[j]), scz.PolynomialCommitment(c, srs_dev[j][0]),
                                         scz.PolynomialCommitment(c, srs_dev[j][1]))
        got = scz.dhyperplonk(c, n, pk, pp).nested()
        comm = c.get_comm()
        c.sync()
        c.close()
        return got, comm

    if net_kind == "local":
        res = LocalTestNet(N, "cuda:0").simulate_network_round(party)
    else:
        import torch
        hub = NativeNcclNet(torch.device("cuda", 0), N)
        res = hub.run_parties(lambda pid, p, net: party(pid, net))
        assert hub.calls["gather"] > 0 and hub.calls["all_gather"] == 1
        hub.close()
    for j in range(N):
        _same_proof(orc, res[j][0], want[j], f"party {j}")
    (gp, gc), (wp, wc, wo) = res[3][0]
    assert all(len(t) == 0 for t in wp[1:]) and len(wp) == 1 + 3 + 3 * (n - 3)     # workers: empty d_ proofs, no leader tail
    assert all(res[j][1] == res[1][1] for j in range(2, N)) and res[0][1][1] > res[1][1][1]   # byte counters by role
    seed_ctx.close()


@pytest.mark.parametrize("variant", ["data_parallel", "permcheck"])
def test_prover_variants_parties_mode(orc, variant):
    """dhyperplonk_data_parallel / dpermcheck with N = 8 real parties (LocalTestNet): every party's output against the
    oracle's 8-party run of the same variant"""
    import scz_b200 as scz
    from oracle import hyperplonk as ohp
    from scz_b200.net import LocalTestNet
    n, l, N = 5, 1, 8
    rng = np.random.default_rng(670)
    opp = orc.pp_new(l)
    seed_ctx = scz.Context(device=0, n_parties=N)
    csz, dsz = ohp.srs_level_sizes(n, l, N)
    dp = variant == "data_parallel"
    opks, srs_dev = [], []
    for j in range(N):
        cdev, csrs = _make_srs(seed_ctx, orc, rng, csz)
        ddev, dsrs = _make_srs(seed_ctx, orc, rng, dsz)
        opks.append(ohp.random_pk(rng, n, l, N, csrs, dsrs, shared=opks[0] if opks else None, data_parallel=dp))
        srs_dev.append((cdev, ddev))
    want = ohp.dhyperplonk(n, opks, opp, orc.PARTIES, N, variant=variant)
    fn = scz.dhyperplonk_data_parallel if dp else scz.dpermcheck

    def party(j, net):
        c = scz.Context(device=0, party_id=j, n_parties=N, net=net)
        pp = scz.PackedSharingParams(c, l)
        pk = scz.PackedProvingParameters(c, n, l, _tables_for_product(opks[j]), scz.PolynomialCommitment(c, srs_dev[j][0]),
                                         scz.PolynomialCommitment(c, srs_dev[j][1]), data_parallel=dp)
        got = fn(c, n, pk, pp).nested()
        c.sync()
        c.close()
        return got

    res = LocalTestNet(N, "cuda:0").simulate_network_round(party)
    for j in range(N):
        _same_proof(orc, res[j], want[j], f"{variant} party {j}")
    seed_ctx.close()


@pytest.mark.parametrize("variant", ["data_parallel", "permcheck"])
def test_prover_variants_leader_mode(orc, variant):
    """dhyperplonk_data_parallel (dhyperplonk.rs:573-960) and dpermcheck (:962-1247): same schedule, s as an input /
    step 2 alone"""
    import scz_b200 as scz
    from oracle import hyperplonk as ohp
    n, l, N = 5, 1, 8
    rng = np.random.default_rng(660)
    ctx = scz.Context(device=0, n_parties=N)
    pp, opp = scz.PackedSharingParams(ctx, l), orc.pp_new(l)
    csz, dsz = ohp.srs_level_sizes(n, l, N)
    cdev, csrs = _make_srs(ctx, orc, rng, csz)
    ddev, dsrs = _make_srs(ctx, orc, rng, dsz)
    dp = variant == "data_parallel"
    opk = ohp.random_pk(rng, n, l, N, csrs, dsrs, data_parallel=dp)
    want = ohp.dhyperplonk(n, [opk], opp, orc.LEADER_SIM, N, variant=variant)[0]
    pk = scz.PackedProvingParameters(ctx, n, l, _tables_for_product(opk), scz.PolynomialCommitment(ctx, cdev),
                                     scz.PolynomialCommitment(ctx, ddev), data_parallel=dp)
    fn = scz.dhyperplonk_data_parallel if dp else scz.dpermcheck
    got = fn(ctx, n, pk, pp).nested()
    _same_proof(orc, got, want, variant)
    if not dp:
        assert got[0] == ([], []) and len(got[1][0]) == 1 + 3 + 3 * (n - 3) + 3
    ctx.close()


def test_local_hyperplonk(orc):
    """the monolithic baseline local_hyperplonk (hyperplonk/src/hyperplonk.rs:15-160), with and without fixed-base tables"""
    import scz_b200 as scz
    from oracle import hyperplonk as ohp
    n = 5
    rng = np.random.default_rng(680)
    ctx = scz.Context(device=0, n_parties=8)
    gc = 1 << n
    pk = {k: orc.random_fr(rng, gc) for k in ("a_evals", "b_evals", "c_evals", "input", "q1", "q2", "eq")}
    pk.update({k: orc.random_fr(rng, 4 * gc) for k in ("m", "ssigma", "sid", "eq_p2")})
    pk.update(challenge=orc.random_fr(rng, n), challengep2=orc.random_fr(rng, n + 2), alpha=orc.random_fr(rng, 1),
              beta=orc.random_fr(rng, 1))
    dev, srs = _make_srs(ctx, orc, rng, [1 << i for i in range(n + 3)])
    pk["commitment"] = srs
    want = ohp.local_hyperplonk(n, pk)
    tabs = {k: v for k, v in pk.items() if k not in ("commitment", "alpha", "beta")}
    tabs["alpha_beta"] = np.concatenate([pk["alpha"], pk["beta"]])
    for pre in (False, True):
        pc = scz.PolynomialCommitment(ctx, dev)
        if pre:
            pc.precompute()
        got = scz.local_hyperplonk(ctx, n, tabs, pc).nested()
        _same_proof(orc, got, want, f"local pre={pre}")
    ctx.close()


def test_dhyperplonk_2p20_full_size_properties(orc):
    """BASELINE config 5 size (2^20 constraints, l = 1, leader mode), checked through size-independent properties:
    the proof is bit-identical with and without the fixed-base tables (different windows, bucket sets and launch
    shapes); every gate sumcheck satisfies the verifier's round identity; the c_open values are mu_0 times the
    evaluations of the share tables (pss2ss in leader mode) computed independently with fix_variable;
    the d_commit of a slice is N times its plain commitment (the leader sums N clones, dpoly_comm.rs:290-292)."""
    import torch
    import scz_b200 as scz
    from oracle import py_twin as tw
    n, l, N = 20, 1, 8
    ctx = scz.Context(device=0, n_parties=N)
    pp = scz.PackedSharingParams(ctx, l)
    pk = scz.PackedProvingParameters.new(ctx, n, l, seed=11, precompute=True)
    proof = scz.dhyperplonk(ctx, n, pk, pp)
    a = [x.clone() for x in (proof.triples, proof.points, proof.values)]
    used = proof.used()
    ctx.msm_use_precompute(False)
    proof2 = scz.dhyperplonk(ctx, n, pk, pp)
    assert proof2.used() == used
    assert torch.equal(a[0][: 3 * used[0]], proof2.triples[: 3 * used[0]]) and torch.equal(a[2][: used[2]], proof2.values[: used[2]])
    p1 = ctx.to_host(ctx.g1_to_affine(a[1][: used[1]].contiguous()))
    p2 = ctx.to_host(ctx.g1_to_affine(proof2.points[: used[1]].contiguous()))
    assert np.array_equal(p1, p2)                                   # same group elements, whatever the Jacobian form
    # and a third time with the batched-affine bucket accumulation switched off (XYZZ mixed additions for every
    # entry, csrc/msm.cu k_msm_accumulate): the first two runs took the affine path for their big sequence
    assert ctx.msm_affine_sequences() >= 2
    ctx.msm_use_precompute(True)
    ctx.msm_set_affine(2)
    seq0 = ctx.msm_affine_sequences()
    proof3 = scz.dhyperplonk(ctx, n, pk, pp)
    assert ctx.msm_affine_sequences() == seq0
    p3 = ctx.to_host(ctx.g1_to_affine(proof3.points[: used[1]].contiguous()))
    assert np.array_equal(p1, p3) and torch.equal(a[0][: 3 * used[0]], proof3.triples[: 3 * used[0]])
    ctx.msm_set_affine(0)
    ctx.msm_use_precompute(False)
    (gp, gc), (wp, wc, wo) = proof2.nested()
    R = tw.R_MOD
    inv2 = pow(2, R - 2, R)
    chi = orc.fr_to_ints(ctx.to_host(pk.t["challenge"]))
    for proof_k in gp:
        tri = [[orc.fr_to_ints(proof_k[i, j:j + 1])[0] for j in range(3)] for i in range(len(proof_k))]
        for i in range(n - 1):
            p0, p1_, p2_ = tri[i]
            r = chi[i]
            val = (p0 * (r - 1) * (r - 2) * inv2 - p1_ * r * (r - 2) + p2_ * r * (r - 1) * inv2) % R
            assert val == (tri[i + 1][0] + tri[i + 1][1]) % R
    for k, name in enumerate(("a_evals", "b_evals", "c_evals")):
        ev = orc.fr_to_ints(ctx.to_host(scz.fix_variable(ctx, pk.t[name], pk.t["challenge"])))[0]
        assert orc.fr_to_ints(gc[k][1][0])[0] == ev * tw.MU0 % R, name
    plain = pk.d_commitment.commit(pk.t["I_p"])
    eight = ctx.g1_mul(plain, ctx.to_device(orc.fr_from_ints([N]), 4))
    assert orc.canon_g1(gc[3][0]) == orc.canon_g1(ctx.to_host(eight))
    ctx.close()


def test_msm_side_stream_same_proof(tmp_path):
    """SCZ_MSM_STREAM=1 (MSM launch sequences on the ctx's low-priority stream, the first one started early under the
    rest of the protocol phase: csrc/msm.cu, Deferred::flush_early) must not change a single byte of the proof.
    The switch is read once per process, so both settings run in fresh interpreters on a high-priority stream."""
    import subprocess
    import sys
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import scz_b200 as scz
torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
ctx = scz.Context(device=0, n_parties=8)
pp = scz.PackedSharingParams(ctx, 1)
out = {}
for n in (6, 12, 16):
    pk = scz.PackedProvingParameters.new(ctx, n, 1, seed=7, precompute=(n == 16))
    for rep in range(2):
        proof = scz.dhyperplonk(ctx, n, pk, pp)
        used = proof.used()
        out[f"t{n}_{rep}"] = ctx.to_host(proof.triples[: used[0]].contiguous())
        out[f"p{n}_{rep}"] = ctx.to_host(ctx.g1_to_affine(proof.points[: used[1]].contiguous()))
        out[f"v{n}_{rep}"] = ctx.to_host(proof.values[: used[2]].contiguous())
np.savez(sys.argv[1], **out)
""" % ROOT
    files = []
    for flag in ("0", "1"):
        f = str(tmp_path / f"proof_{flag}.npz")
        env = dict(os.environ, SCZ_MSM_STREAM=flag)
        r = subprocess.run([sys.executable, "-c", code, f], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        files.append(np.load(f))
    a, b = files
    assert sorted(a.files) == sorted(b.files) and len(a.files) == 18
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k
    for n in (6, 12, 16):   # and a proof repeats itself
        for kind in "tpv":
            assert np.array_equal(b[f"{kind}{n}_0"], b[f"{kind}{n}_1"]), (kind, n)


def test_proof_reader_pipelined_readback(orc):
    """ProofReader (api.py): proofs read back one behind the prover on a separate stream are the same proofs as the
    synchronous `to_host()` gives (field elements byte for byte; points as group elements -- the Jacobian representative
    depends on the order in which the counting sort's atomics lay out a bucket's entries), and the status snapshot travels with each proof (a division by zero in proof 1 is reported
    for proof 1 only; arkworks panics there, dhyperplonk.rs:338-339)."""
    import torch

    import scz_b200 as scz
    ctx = scz.Context(device=0, n_parties=8)
    pp = scz.PackedSharingParams(ctx, 1)
    n = 8
    pks = [scz.PackedProvingParameters.new(ctx, n, 1, seed=11 + i) for i in range(3)]
    want = [scz.dhyperplonk(ctx, n, pk, pp).to_host() for pk in pks]
    reader = scz.ProofReader(ctx, depth=2)
    got, pending = [], None
    for pk in pks:
        ticket = scz.dhyperplonk(ctx, n, pk, pp).to_host_async(reader)
        if pending is not None:
            got.append(reader.collect(pending))
        pending = ticket
    got.append(reader.collect(pending))
    for a, b in zip(got, want):
        assert a[0].shape == b[0].shape and np.array_equal(a[0], b[0])          # triples
        assert a[1].shape == b[1].shape and orc.canon_g1(a[1]) == orc.canon_g1(b[1])   # points
        assert a[2].shape == b[2].shape and np.array_equal(a[2], b[2])          # values
    # a zero denominator in the middle proof: only that ticket raises
    a = ctx.to_device(np.ones((4, 4), dtype=np.uint64), 4)
    z = torch.zeros_like(a)
    t0 = scz.dhyperplonk(ctx, n, pks[0], pp).to_host_async(reader)
    scz.fr_pointwise(ctx, "div", a, z)                      # raises the sticky bit in stream order
    t1 = scz.dhyperplonk(ctx, n, pks[1], pp).to_host_async(reader)
    reader.collect(t0)
    with pytest.raises(ZeroDivisionError):
        reader.collect(t1)
    assert ctx.take_status() == 0
