"""Committed fixture tests/golden/golden_v1.json (made by tests/golden/make_golden.py from the big-integer twin, the
reference's three known answers and public BLS12-381 constants) against
  * the C oracle                      (CPU, every run), and
  * the CUDA library through its C ABI (-m gpu).
The reference has no golden vectors of its own for this path."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_v1.json")))
T = GOLD["twin"]


def ints(xs):
    return [int(x, 16) for x in xs]


def pts(ps):
    return [(0, 0, 1) if p is None else (int(p[0], 16), int(p[1], 16), 0) for p in ps]


def aff13(orc, ps):
    """golden affine points -> oracle affine rows (x | y Montgomery, infinity word)"""
    out = np.zeros((len(ps), 13), dtype=np.uint64)
    for i, p in enumerate(ps):
        if p is None:
            out[i, 12] = 1
        else:
            out[i, 0:6] = orc.fq_from_ints([int(p[0], 16)])[0]
            out[i, 6:12] = orc.fq_from_ints([int(p[1], 16)])[0]
    return out


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_fixture_is_reproducible(tmp_path):
    """the committed file is what the committed script writes"""
    import subprocess
    import sys
    src = os.path.join(HERE, "golden", "make_golden.py")
    code = open(src).read().replace('os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json")', repr(str(tmp_path / "g.json")))
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE))
    p = tmp_path / "mk.py"
    p.write_text(code.replace("ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))", f"ROOT = {os.path.dirname(HERE)!r}"))
    subprocess.run([sys.executable, str(p)], check=True, env=env, capture_output=True)
    assert json.load(open(tmp_path / "g.json")) == GOLD


def test_oracle_reference_known_answers(orc):
    ka = GOLD["ref_known_answers"]
    assert list(orc.sub_index(26)) == ka["sub_index_26"]                      # dacc_product.rs:442-448
    vx0, vx1, v1x = orc.acc_product(orc.fr_from_ints([1, 2, 3, 4]))           # dacc_product.rs:450-466
    assert [orc.fr_to_ints(v) for v in (vx0, vx1, v1x)] == ka["acc_product_1_2_3_4"]


def test_oracle_public_constants(orc):
    pub = GOLD["public"]
    from oracle import py_twin as tw
    assert int(pub["r"], 16) == tw.R_MOD and int(pub["p"], 16) == tw.P_MOD
    gen = orc.canon_g1(orc.g1_from_affine(orc.g1_generator()))
    assert gen == pts([pub["g1_generator"]])
    pp = orc.pp_new(1)
    one = orc.fr_from_ints([1])
    assert orc.fr_to_ints(orc.pack_from_public(pp, one))[0] == int(pub["lambda0_4_over_7"], 16)
    assert orc.fr_to_ints(orc.pack_single(pp, one))[0] == int(pub["mu0_pss2ss"], 16)


def test_oracle_fields(orc):
    for name, frm, to, ops in (("fr", orc.fr_from_ints, orc.fr_to_ints, (orc.fr_mul, orc.fr_add, orc.fr_sub)),
                               ("fq", orc.fq_from_ints, orc.fq_to_ints, (orc.fq_mul, orc.fq_add, orc.fq_sub))):
        a, b = frm(ints(T[name]["a"])), frm(ints(T[name]["b"]))
        for op, fn in zip(("mul", "add", "sub"), ops):
            assert to(fn(a, b)) == ints(T[name][op]), (name, op)
    a = ints(T["fr"]["a"])
    nz = [i for i, v in enumerate(a) if v]
    inv = orc.fr_to_ints(orc.fr_inv(orc.fr_from_ints([a[i] for i in nz])))
    assert inv == [ints(T["fr"]["inv_a"])[i] for i in nz]


def test_oracle_g1(orc):
    g = T["g1"]
    k = orc.fr_from_ints(ints(g["k"]))
    P = orc.g1_from_affine(orc.g1_gen_mul(k))
    assert orc.canon_g1(P) == pts(g["k_times_generator"])
    assert orc.canon_g1(orc.g1_add(P, np.roll(P, -1, axis=0))) == pts(g["sum_with_next"])
    assert orc.canon_g1(orc.g1_double(P)) == pts(g["doubled"])


def test_oracle_msm_and_d_msm(orc):
    m = T["msm"]
    bases = aff13(orc, m["bases"])
    assert orc.canon_g1(orc.g1_from_affine(orc.g1_gen_mul(orc.fr_from_ints(ints(m["base_scalars"]))))) == pts(m["bases"])
    sc = orc.fr_from_ints(ints(m["scalars"]))
    for algo in ("ark", "naive"):
        assert orc.canon_g1(orc.msm(bases, sc, algo=algo)) == pts([m["msm"]])
    out = orc.d_msm(orc.pp_new(1), orc.LEADER_SIM, [[bases]], [[sc]])
    assert orc.canon_g1(out[0]) == pts([m["d_msm_leader_l1"]])


def test_oracle_pss(orc):
    for e in T["pss"]:
        l = e["l"]
        pp = orc.pp_new(l)
        sh = orc.pack_from_public(pp, orc.fr_from_ints(ints(e["secrets"])))
        assert orc.fr_to_ints(sh) == ints(e["pack_from_public"])
        assert orc.fr_to_ints(orc.unpack(pp, sh)) == ints(e["unpack"]) == ints(e["secrets"])
        assert orc.fr_to_ints(orc.pack_single(pp, orc.fr_from_ints(ints(e["secrets"][:1])))) == ints(e["pack_single"])
        prod = orc.fr_mul(sh, orc.fr_from_ints(ints(e["other_shares"])))
        assert orc.fr_to_ints(orc.unpack2(pp, prod)) == ints(e["unpack2_of_product"])
        if "g1_secrets" in e:
            gs = orc.g1_from_affine(aff13(orc, e["g1_secrets"]))
            gsh = orc.pack_from_public(pp, gs, kind=1)
            assert orc.canon_g1(gsh) == pts(e["g1_pack_from_public"])
            assert orc.canon_g1(orc.unpack(pp, gsh, kind=1)) == pts(e["g1_unpack"]) == pts(e["g1_secrets"])


def test_oracle_sumcheck_mle_tree(orc):
    s = T["sumcheck_product"]
    proof = orc.sumcheck_product(orc.fr_from_ints(ints(s["f"])), orc.fr_from_ints(ints(s["g"])), orc.fr_from_ints(ints(s["challenge"])))
    assert [orc.fr_to_ints(np.asarray(tr).reshape(3, 4)) for tr in np.asarray(proof).reshape(-1, 3, 4)] == [ints(tr) for tr in s["proof"]]
    m = T["mle"]
    assert orc.fr_to_ints(orc.fix_variable(orc.fr_from_ints(ints(m["evals"])), orc.fr_from_ints(ints(m["point"])))) == [int(m["value"], 16)]
    a = T["acc_product"]
    vx0, vx1, v1x = orc.acc_product(orc.fr_from_ints(ints(a["x"])))
    assert (orc.fr_to_ints(vx0), orc.fr_to_ints(vx1), orc.fr_to_ints(v1x)) == (ints(a["vx0"]), ints(a["vx1"]), ints(a["v1x"]))


# ------------------------------------------------------------------------------------------------ GPU: the C ABI
@pytest.fixture(scope="module")
def ctx():
    import scz_b200 as scz
    c = scz.Context(device=0, n_parties=8)
    yield c
    c.close()


@pytest.mark.gpu
def test_cuda_fields_and_g1(orc, ctx):
    for name, cols, frm, to in (("fr", 4, orc.fr_from_ints, orc.fr_to_ints), ("fq", 6, orc.fq_from_ints, orc.fq_to_ints)):
        a, b = ctx.to_device(frm(ints(T[name]["a"])), cols), ctx.to_device(frm(ints(T[name]["b"])), cols)
        fn = ctx.fr_op if name == "fr" else ctx.fq_op
        for op in ("mul", "add", "sub"):
            assert to(ctx.to_host(fn(op, a, b))) == ints(T[name][op]), (name, op)
    g = T["g1"]
    k = ctx.to_device(orc.fr_from_ints(ints(g["k"])), 4)
    aff = ctx.g1_generator_mul(k)                                   # packed affine, infinity = all zero
    from tests.gpu_util import oracle_affine
    jac = ctx.to_device(orc.g1_from_affine(oracle_affine(ctx.to_host(aff))), 18)
    assert orc.canon_g1(ctx.to_host(jac)) == pts(g["k_times_generator"])
    assert orc.canon_g1(ctx.to_host(ctx.g1_add(jac, torch_roll(jac)))) == pts(g["sum_with_next"])
    assert orc.canon_g1(ctx.to_host(ctx.g1_double(jac))) == pts(g["doubled"])
    enc = ctx.g1_serialize_compressed(jac).cpu().numpy()
    assert [bytes(r).hex() for r in enc] == g["compressed"]
    pub = GOLD["public"]
    assert bytes(enc[1]).hex() == pub["g1_generator_compressed"] and bytes(enc[0]).hex() == pub["g1_infinity_compressed"]
    assert bytes(enc[4]).hex() == pub["g1_neg_generator_compressed"]          # k = r - 1


def torch_roll(t):
    import torch
    return torch.roll(t, -1, 0).contiguous()


@pytest.mark.gpu
def test_cuda_msm_and_d_msm(orc, ctx):
    import scz_b200 as scz
    from tests.gpu_util import packed_affine
    m = T["msm"]
    bases = packed_affine(aff13(orc, m["bases"]))
    sc = orc.fr_from_ints(ints(m["scalars"]))
    assert orc.canon_g1(scz.msm(ctx, bases, sc)) == pts([m["msm"]])
    out = scz.d_msm(ctx, scz.PackedSharingParams(ctx, 1), [bases], [sc])
    assert orc.canon_g1(out) == pts([m["d_msm_leader_l1"]])


@pytest.mark.gpu
def test_cuda_pss(orc):
    import scz_b200 as scz
    for e in T["pss"]:
        l = e["l"]
        c = scz.Context(device=0, n_parties=8 * l)
        pp = scz.PackedSharingParams(c, l)
        sh = pp.pack_from_public(orc.fr_from_ints(ints(e["secrets"])))
        assert orc.fr_to_ints(sh.reshape(-1, 4)) == ints(e["pack_from_public"])
        assert orc.fr_to_ints(pp.unpack(sh.reshape(-1, 4)).reshape(-1, 4)) == ints(e["secrets"])
        assert orc.fr_to_ints(pp.pack_single(orc.fr_from_ints(ints(e["secrets"][:1]))).reshape(-1, 4)) == ints(e["pack_single"])
        prod = orc.fr_mul(sh.reshape(-1, 4), orc.fr_from_ints(ints(e["other_shares"])))
        assert orc.fr_to_ints(pp.unpack2(prod).reshape(-1, 4)) == ints(e["unpack2_of_product"])
        if "g1_secrets" in e:
            gs = orc.g1_from_affine(aff13(orc, e["g1_secrets"]))
            gsh = pp.pack_from_public(gs, kind="g1")
            assert orc.canon_g1(gsh.reshape(-1, 18)) == pts(e["g1_pack_from_public"])
            assert orc.canon_g1(pp.unpack(gsh.reshape(-1, 18), kind="g1").reshape(-1, 18)) == pts(e["g1_secrets"])
        c.close()


@pytest.mark.gpu
def test_cuda_sumcheck_mle_tree(orc, ctx):
    import scz_b200 as scz
    s = T["sumcheck_product"]
    proof = scz.sumcheck_product(ctx, orc.fr_from_ints(ints(s["f"])), orc.fr_from_ints(ints(s["g"])), orc.fr_from_ints(ints(s["challenge"])))
    got = orc.fr_to_ints(np.asarray(proof).reshape(-1, 4))
    assert got == [v for tr in s["proof"] for v in ints(tr)]
    m = T["mle"]
    val = scz.fix_variable(ctx, orc.fr_from_ints(ints(m["evals"])), orc.fr_from_ints(ints(m["point"])))
    assert orc.fr_to_ints(np.asarray(val).reshape(-1, 4)) == [int(m["value"], 16)]
    a = T["acc_product"]
    tree = np.asarray(scz.acc_product_tree(ctx, orc.fr_from_ints(ints(a["x"])))).reshape(-1, 4)
    t = orc.fr_to_ints(tree)                     # the 2m-entry table of dacc_product.rs:30-57
    assert (t[0::2], t[1::2], t[len(t) // 2:]) == (ints(a["vx0"]), ints(a["vx1"]), ints(a["v1x"]))
    ka = GOLD["ref_known_answers"]["acc_product_1_2_3_4"]
    t = orc.fr_to_ints(np.asarray(scz.acc_product_tree(ctx, orc.fr_from_ints([1, 2, 3, 4]))).reshape(-1, 4))
    assert [t[0::2], t[1::2], t[len(t) // 2:]] == ka
