"""Generates tests/golden/golden_v1.json (committed).  Run from the repo root:  python tests/golden/make_golden.py

The reference holds NO byte-level golden vectors for this path: its tests are properties, plus
three tiny known answers (dacc_product.rs:442-466, operator.rs:42-49).  This fixture therefore carries
  * the reference's own known answers, restated (`ref_known_answers`);
  * public constants of BLS12-381 and of ark-bls12-381's wire format that anyone can check against the published
    curve parameters (`public`);
  * input / output vectors for every layer of the path computed by the BIG-INTEGER twin (oracle/py_twin.py: plain
    Python integers, schoolbook group law, O(n^2) DFTs -- it shares no code and no algorithm with the C oracle or
    the CUDA library) on seeded inputs (`twin`).
It pins the C oracle (tests/test_golden.py, CPU) and the CUDA path (same file, -m gpu) to a third implementation,
not to arkworks itself: parity with the Rust reference stays "unpinned" in the sense of DESIGN.md section 5.
All integers are canonical (non-Montgomery) values as hex strings; G1 points are affine [x, y] or null."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import py_twin as tw  # noqa: E402

R, P = tw.R_MOD, tw.P_MOD
G = (tw.G1_X, tw.G1_Y)


def hx(v):
    return hex(v)


def pt(p):
    return None if p is tw.INF else [hx(p[0]), hx(p[1])]


def main():
    rnd = random.Random(0x5CA1AB1E)
    fr = lambda: rnd.randrange(R)          # noqa: E731
    fq = lambda: rnd.randrange(P)          # noqa: E731
    out = {"version": 1, "generator": "tests/golden/make_golden.py (oracle/py_twin.py big-integer twin)"}

    out["ref_known_answers"] = {
        "sub_index_26": [20, 21],                                          # dacc_product.rs:442-448
        "acc_product_1_2_3_4": [[1, 3, 2, 24], [2, 4, 12, 0], [2, 12, 24, 0]],   # dacc_product.rs:450-466
        "transpose_2x3": [[[1, 2, 3], [4, 5, 6]], [[1, 4], [2, 5], [3, 6]]],     # operator.rs:42-49
    }
    out["public"] = {
        "r": hx(R), "p": hx(P), "g1_generator": pt(G), "fr_generator": 7, "two_adicity": 32,
        # Zcash / IETF compressed encodings (48 bytes, big endian x, flag bits in the top byte)
        "g1_generator_compressed": tw.g1_serialize_compressed(G).hex(),
        "g1_neg_generator_compressed": tw.g1_serialize_compressed(tw.g1_neg(G)).hex(),
        "g1_infinity_compressed": tw.g1_serialize_compressed(tw.INF).hex(),
        # l = 1 constants of BASELINE.md section 4
        "lambda0_4_over_7": hx(tw.LAMBDA0), "mu0_pss2ss": hx(tw.MU0), "omega8": hx(tw.OMEGA8),
    }

    t = {}
    edge_r = [0, 1, 2, R - 1, R - 2, (R - 1) // 2, (1 << 32) - 1, 1 << 64, 1 << 200]
    a = edge_r + [fr() for _ in range(23)]
    b = list(reversed(edge_r)) + [fr() for _ in range(23)]
    t["fr"] = {"a": [hx(x) for x in a], "b": [hx(x) for x in b],
               "mul": [hx(x * y % R) for x, y in zip(a, b)], "add": [hx((x + y) % R) for x, y in zip(a, b)],
               "sub": [hx((x - y) % R) for x, y in zip(a, b)],
               "inv_a": [hx(tw.finv(x, R)) if x else hx(0) for x in a]}
    edge_p = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (1 << 64) - 1, 1 << 380, (1 << 381) - 1 - (1 << 200)]
    a = edge_p + [fq() for _ in range(23)]
    b = list(reversed(edge_p)) + [fq() for _ in range(23)]
    t["fq"] = {"a": [hx(x) for x in a], "b": [hx(x) for x in b],
               "mul": [hx(x * y % P) for x, y in zip(a, b)], "add": [hx((x + y) % P) for x, y in zip(a, b)],
               "sub": [hx((x - y) % P) for x, y in zip(a, b)]}

    ks = [0, 1, 2, 3, R - 1, R - 2, (R + 1) // 2] + [fr() for _ in range(9)]
    pts = [tw.g1_mul(G, k) for k in ks]
    t["g1"] = {"k": [hx(k) for k in ks], "k_times_generator": [pt(q) for q in pts],
               "sum_with_next": [pt(tw.g1_add(pts[i], pts[(i + 1) % len(pts)])) for i in range(len(pts))],
               "doubled": [pt(tw.g1_add(q, q)) for q in pts],
               "compressed": [tw.g1_serialize_compressed(q).hex() for q in pts]}

    # MSM (dmsm.rs:23) and the leader-mode d_msm (dmsm.rs:9-43) for l = 1: (4/7) * msm
    m = 48
    bk = [fr() for _ in range(m)]
    sc = [fr() for _ in range(m - 4)] + [0, 1, R - 1, 2]
    bases = [tw.g1_mul(G, k) for k in bk]
    msm = tw.g1_msm(bases, sc)
    t["msm"] = {"base_scalars": [hx(k) for k in bk], "bases": [pt(q) for q in bases], "scalars": [hx(s) for s in sc],
                "msm": pt(msm), "d_msm_leader_l1": pt(tw.g1_mul(msm, tw.LAMBDA0))}

    # PSS maps (pss.rs:38-171) over Fr for l = 1, 2, 4 and over G1 for l = 1, 2
    t["pss"] = []
    for l in (1, 2, 4):
        ps = tw.PSS(l)
        secrets = [fr() for _ in range(l)]
        shares = ps.pack_from_public(secrets)
        other = ps.pack_from_public([fr() for _ in range(l)])
        prod = [x * y % R for x, y in zip(shares, other)]
        e = {"l": l, "secrets": [hx(x) for x in secrets], "pack_from_public": [hx(x) for x in shares],
             "unpack": [hx(x) for x in ps.unpack(shares)], "pack_single": [hx(x) for x in ps.pack_single(secrets[0])],
             "other_shares": [hx(x) for x in other], "unpack2_of_product": [hx(x) for x in ps.unpack2(prod)]}
        if l <= 2:
            gs = [tw.g1_mul(G, s) for s in secrets]
            gsh = ps.pack_from_public(gs, "g1")
            e["g1_secrets"] = [pt(q) for q in gs]
            e["g1_pack_from_public"] = [pt(q) for q in gsh]
            e["g1_unpack"] = [pt(q) for q in ps.unpack(gsh, "g1")]
        t["pss"].append(e)

    # product sumcheck (dsumcheck.rs:28-90), fix_variable / MLE evaluation (mle.rs:88-104), product tree (dacc_product.rs:30-57)
    n = 4
    f = [fr() for _ in range(1 << n)]
    g = [fr() for _ in range(1 << n)]
    ch = [fr() for _ in range(n)]
    proof = tw.sumcheck_product(f, g, ch)
    assert tw.check_sumcheck_product(sum(x * y for x, y in zip(f, g)) % R, proof, ch, n)
    t["sumcheck_product"] = {"f": [hx(x) for x in f], "g": [hx(x) for x in g], "challenge": [hx(x) for x in ch],
                             "proof": [[hx(x) for x in tr] for tr in proof]}
    t["mle"] = {"evals": [hx(x) for x in f], "point": [hx(x) for x in ch], "value": hx(tw.mle_eval(f, ch))}
    x = [fr() for _ in range(8)]
    vx0, vx1, v1x = tw.acc_product(x)
    t["acc_product"] = {"x": [hx(v) for v in x], "vx0": [hx(v) for v in vx0], "vx1": [hx(v) for v in vx1], "v1x": [hx(v) for v in v1x]}
    out["twin"] = t

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
        fh.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
