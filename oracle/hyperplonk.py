"""ORACLE (test infrastructure) -- CPU restatement of the collaborative HyperPlonk prover,
`dhyperplonk` (hyperplonk/src/dhyperplonk.rs:159-571), composed from the oracle's C primitives
(oracle/src/*.c through oracle/oracle.py).  Only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import this module.  PARITY UNPINNED at the arkworks byte boundary (oracle/src/oracle.h).

The reference draws three vectors from entropy inside the function (`local_s_p`, `local_s`, `eq`,
dhyperplonk.rs:188-190); here (and in the product's API) they are explicit inputs so that runs can be
compared.  Every statement below cites the reference line it restates.

A "pk" is a dict of numpy arrays (Fr: (len, 4) uint64 Montgomery) mirroring PackedProvingParameters
(dhyperplonk.rs:22-62) -- only the fields `dhyperplonk` reads -- plus:
  'c_commitment', 'd_commitment': oracle.Srs      (dhyperplonk.rs:53-54)
  'alpha', 'beta': (1, 4)                          (:50-51)
  'local_s_p', 'local_s', 'eq_leader': injected entropy (:188-190)
"""
import numpy as np

from . import oracle as orc

PK_TABLES = ("V", "a_evals", "b_evals", "c_evals", "I", "S1", "S2", "I_p", "S1_p", "S2_p", "ssigma_p", "sid_p", "eq",
             "eq_r1_p", "eq_r2_p", "challenge", "challenge_r1", "challenge_r2", "alpha", "beta", "local_s_p", "local_s",
             "eq_leader")


def table_sizes(n, l, N):
    """lengths of every table `dhyperplonk` reads, PackedProvingParameters::new (dhyperplonk.rs:65-157) + :188-190"""
    gc = 1 << n
    return {
        "V": gc * 4 // l, "a_evals": gc // l, "b_evals": gc // l, "c_evals": gc // l,   # :70-73 (fix_variable drops 2 vars)
        "I": gc // l, "S1": gc // l, "S2": gc // l,                                      # :75, :79-80
        "I_p": gc // N, "S1_p": gc // N, "S2_p": gc // N,                                # :77, :81-82
        "ssigma_p": gc * 4 // N, "sid_p": gc * 4 // N,                                   # :85, :90
        "eq": gc // l, "eq_r1_p": gc * 4 // N, "eq_r2_p": gc * 4 // N,                   # :93, :96, :98
        "challenge": n, "challenge_r1": n + 2, "challenge_r2": n + 2,                    # :103-105
        "alpha": 1, "beta": 1,                                                           # :107-108
        "local_s_p": gc * 4 // N, "local_s": gc * 4 // N // l, "eq_leader": 8 * l,       # :188-190 (pp.n = 8 l)
    }


def srs_level_sizes(n, l, N):
    """(c_commitment, d_commitment) level lengths: new_single(n+2, pp) :196-219, new_random(n+2, N) :220-233"""
    c = [max(1, (1 << i) // l) for i in range(n + 3)]
    d = [1 << i for i in range(n + 2 - (N.bit_length() - 1) + 1)]
    return c, d


def _rep(x, m):
    return np.repeat(np.ascontiguousarray(x, dtype=np.uint64).reshape(1, 4), m, axis=0)


def dhyperplonk(n, pks, pp, mode, N, algo="ark", variant="full"):
    """pks: one pk per party (PARTIES: N of them, LEADER_SIM: the leader's alone).
    Returns one result dict per party:
      gate_identity_proofs       list of (cnt, 3, 4)
      gate_identity_commitments  list of (com (1,18), (value (1,4), proofs (k,18)))
      wiring_proofs, wiring_commits, wiring_opens    as the reference's tuple (dhyperplonk.rs:567-570);
    entries a non-leader gets empty from d_sumcheck_product / d_open are empty arrays there.
    variant: "full" = dhyperplonk (:159-571); "data_parallel" = dhyperplonk_data_parallel (:573-960: pk["local_s"] is the
    whole s, 4 * 2^n / l entries, :603; no exchange); "permcheck" = dpermcheck (:962-1247: step 2 alone)."""
    assert variant in ("full", "data_parallel", "permcheck")
    gate = variant != "permcheck"
    P = N if mode == orc.PARTIES else 1
    assert len(pks) == P
    csrs = [pk["c_commitment"] for pk in pks]
    dsrs = [pk["d_commitment"] for pk in pks]
    res = [dict(gate_identity_proofs=[], gate_identity_commitments=[], wiring_proofs=[], wiring_commits=[],
                wiring_opens=[]) for _ in range(P)]
    col = lambda name: [pk[name] for pk in pks]   # noqa: E731

    def c_commit1(name_or_tabs):
        tabs = col(name_or_tabs) if isinstance(name_or_tabs, str) else name_or_tabs
        out = orc.c_commit(csrs, pp, mode, [[t] for t in tabs], algo)      # (P, 1, 18)
        return [out[j, 0:1].copy() for j in range(P)]

    def d_commit(tabs):
        com = orc.d_commit(dsrs, mode, N, tabs, algo)                      # every party gets the sum (dpoly_comm.rs:290-292)
        return [com.copy() for _ in range(P)]

    def c_sumcheck(f, g, ch):
        out = orc.c_sumcheck_product(pp, mode, f, g, ch)
        return [out[j].copy() for j in range(P)]

    def d_sumcheck(f, g, ch):
        lead = orc.d_sumcheck_product(mode, N, f, g, ch)
        return [lead] + [np.zeros((0, 3, 4), dtype=np.uint64) for _ in range(P - 1)]   # dsumcheck.rs:507-509

    def c_open(tabs, point):
        val, proofs = orc.c_open(csrs, pp, mode, tabs, point, algo)
        return [(val[j:j + 1].copy(), proofs[j].copy()) for j in range(P)]

    def d_open(tabs, point):
        val, proofs = orc.d_open(dsrs, mode, N, tabs, point, algo)
        empty = (np.zeros((1, 4), dtype=np.uint64), np.zeros((0, 18), dtype=np.uint64))   # dpoly_comm.rs:387
        return [(val, proofs)] + [empty for _ in range(P - 1)]

    def push(key, per_party):
        for j in range(P):
            res[j][key].append(per_party[j])

    # Step 1: commit (:196-217)
    if gate:
        com_a, com_b, com_c = c_commit1("a_evals"), c_commit1("b_evals"), c_commit1("c_evals")
        com_I, com_S1, com_S2 = d_commit(col("I_p")), d_commit(col("S1_p")), d_commit(col("S2_p"))

    # Step 3: gate identity (:222-260)
    ch = pks[0]["challenge"]
    if gate:
        push("gate_identity_proofs", c_sumcheck(col("eq"), col("S1"), ch))                          # :230-231
        sum_ab = [orc.fr_add(pk["a_evals"], pk["b_evals"]) for pk in pks]                           # :233-238
        push("gate_identity_proofs", c_sumcheck(col("S1"), sum_ab, ch))                             # :240-241
        push("gate_identity_proofs", c_sumcheck(col("eq"), col("S2"), ch))                          # :243-244
        push("gate_identity_proofs", c_sumcheck(col("a_evals"), col("b_evals"), ch))                # :245-246
        push("gate_identity_proofs", c_sumcheck(col("S2"), col("a_evals"), ch))                     # :247-248
        sum_ci = [orc.fr_sub(pk["I"], pk["c_evals"]) for pk in pks]                                 # :251-256  -c + I
        push("gate_identity_proofs", c_sumcheck(col("eq"), sum_ci, ch))                             # :258-259

    # Step 2: wiring identity (:263-513)
    # 2.a (:270-294): N hub rounds, hub i sends its local_s to everyone -> s = local_s^(0) | ... | local_s^(N-1);
    # without `comm` the own vector is appended N times (:289-293)
    if variant == "data_parallel":                       # s drawn locally by every party (:603)
        s = col("local_s")
    elif mode == orc.PARTIES:
        s_all = np.concatenate([pk["local_s"] for pk in pks])
        s = [s_all for _ in range(P)]
    else:
        s = [np.concatenate([pks[0]["local_s"]] * N)]
    local_s_p = col("local_s_p")
    push("wiring_commits", d_commit(local_s_p))                                                     # 2.b :297-302
    r1, r2 = pks[0]["challenge_r1"], pks[0]["challenge_r2"]
    push("wiring_proofs", c_sumcheck(s, col("V"), r1))                                              # 2.c :304
    push("wiring_opens", c_open(col("V"), r1))                                                      # 2.d :306-320
    push("wiring_opens", c_open(col("V"), r2))
    push("wiring_opens", d_open(local_s_p, r2))
    # 2.e (:324-340)
    num, den, h_p = [], [], []
    for pk in pks:
        m = len(pk["local_s_p"])
        alpha, beta = _rep(pk["alpha"], m), _rep(pk["beta"], m)
        nu = orc.fr_add(orc.fr_add(pk["local_s_p"], orc.fr_mul(alpha, pk["sid_p"])), beta)          # :326-331
        de = orc.fr_add(orc.fr_add(pk["eq_r1_p"], orc.fr_mul(alpha, pk["ssigma_p"])), beta)         # :332-337
        num.append(nu)
        den.append(de)
        h_p.append(orc.fr_mul(nu, orc.fr_inv(de)))                                                  # :339
    subtrees, top = orc.d_acc_product(mode, N, h_p)                                                 # :342
    v1x = [t[len(t) // 2:].copy() for t in subtrees]                                                # :344-348
    vx0 = [t[0::2].copy() for t in subtrees]                                                        # :349-353
    vx1 = [t[1::2].copy() for t in subtrees]                                                        # :354-359
    for tabs in (col("ssigma_p"), col("sid_p"), h_p, num, den, v1x, vx0, vx1):                      # :363-380
        push("wiring_commits", d_commit(tabs))
    for tabs in (col("ssigma_p"), col("sid_p"), h_p, num, den):                                     # :383-407
        push("wiring_opens", d_open(tabs, r2))
    eq_r2_p = col("eq_r2_p")
    push("wiring_proofs", d_sumcheck(den, eq_r2_p, r2))                                             # :411-413
    push("wiring_proofs", d_sumcheck(h_p, den, r2))
    push("wiring_proofs", d_sumcheck(num, eq_r2_p, r2))
    sN = N.bit_length() - 1                                                                         # :418
    cur_v1x = [t[:len(t) // 2] for t in v1x]                                                        # :419-422
    cur_vx0 = [t[:len(t) // 2] for t in vx0]
    cur_vx1 = [t[:len(t) // 2] for t in vx1]
    cur_eq = [t[:len(t) // 2] for t in eq_r2_p]
    for i in range(1, n - sN + 1):                                                                  # :423-478
        chi = r2[i:]
        push("wiring_proofs", d_sumcheck(cur_eq, cur_v1x, chi))
        push("wiring_proofs", d_sumcheck(cur_eq, cur_vx0, chi))
        push("wiring_proofs", d_sumcheck(cur_vx0, cur_vx1, chi))
        push("wiring_opens", d_open(cur_v1x, chi))
        push("wiring_opens", d_open(cur_vx0, chi))
        push("wiring_opens", d_open(cur_vx1, chi))
        cur_v1x = [t[len(t) // 2:] for t in cur_v1x]
        cur_vx0 = [t[len(t) // 2:] for t in cur_vx0]
        cur_vx1 = [t[len(t) // 2:] for t in cur_vx1]
        cur_eq = [t[len(t) // 2:] for t in cur_eq]
    # leader tree (:480-511): party 0 only
    lt = top
    lv1x, lvx0, lvx1 = lt[len(lt) // 2:].copy(), lt[0::2].copy(), lt[1::2].copy()
    pt = r2[:sN]
    d0 = dsrs[0]
    for t in (lvx0, lvx1, lv1x):                                                                    # :500-505
        res[0]["wiring_commits"].append(orc.commit(d0, t, algo))
        res[0]["wiring_opens"].append(orc.open_(d0, t, pt, algo))
    eql = pks[0]["eq_leader"]
    res[0]["wiring_proofs"].append(orc.sumcheck_product(eql, lv1x, pt))                             # :507-509
    res[0]["wiring_proofs"].append(orc.sumcheck_product(eql, lvx0, pt))
    res[0]["wiring_proofs"].append(orc.sumcheck_product(lvx0, lvx1, pt))

    # Open (:517-554)
    if not gate:
        return res
    for com, name in ((com_a, "a_evals"), (com_b, "b_evals"), (com_c, "c_evals")):
        op = c_open(col(name), ch)
        push("gate_identity_commitments", [(com[j], op[j]) for j in range(P)])
    for com, name in ((com_I, "I_p"), (com_S1, "S1_p"), (com_S2, "S2_p")):
        op = d_open(col(name), ch)
        push("gate_identity_commitments", [(com[j], op[j]) for j in range(P)])
    return res


def random_pk(rng, n, l, N, srs_c=None, srs_d=None, shared=None, data_parallel=False):
    """PackedProvingParameters::new with numpy randomness (tables only; SRS passed in).  `shared`: a pk whose
    public values (challenges, alpha, beta) are reused -- all parties must agree on them."""
    pk = {}
    sizes = table_sizes(n, l, N)
    if data_parallel:
        sizes["local_s"] = (1 << n) * 4 // l             # dhyperplonk_data_parallel: the whole s (:603)
    for name, ln in sizes.items():
        pk[name] = orc.random_fr(rng, ln)
    if shared is not None:
        for name in ("challenge", "challenge_r1", "challenge_r2", "alpha", "beta"):
            pk[name] = shared[name]
    pk["c_commitment"], pk["d_commitment"] = srs_c, srs_d
    return pk


# ---------------------------------------------------------------------------------------------------------------
# The collaborative (PSS) permutation check: hyperplonk/src/dhyperplonk.rs:1249-1385 and its masked product
# accumulation dist-primitive/src/dacc_product.rs:66-363.  Restated statement by statement, including what the build
# without `comm` substitutes for received data; the reference itself says "We do not guarantee correctness here"
# (dacc_product.rs:353).
def _pack_chunks(pp, vals):
    """vals.chunks(l).map(pack_from_public) then transpose (dacc_product.rs:121-130): -> (n, len(vals) / l, 4)"""
    l = pp.l
    packs = [orc.pack_from_public(pp, vals[c:c + l]) for c in range(0, len(vals), l)]
    if not packs:
        return np.zeros((pp.n, 0, 4), dtype=np.uint64)
    return np.stack(packs, axis=1)


def _merge(rows):
    """merge (dacc_product.rs:416-428)"""
    r = len(rows[0])
    num = 1
    while num < r + 1:
        num <<= 1
    num >>= 1
    out, start = [], 0
    while num and start + num <= r:
        for row in rows:
            out.append(row[start:start + num])
        start += num
        num >>= 1
    return np.concatenate(out) if out else np.zeros((0, 4), dtype=np.uint64)


def c_acc_product(pp, mode, inputs):
    """dacc_product.rs:296-363 -> (subtrees per party, leader tree of N * N entries)"""
    N = pp.n
    P = N if mode == orc.PARTIES else 1
    subtrees = [orc.acc_product_tree(x) for x in inputs]
    assert 2 * len(inputs[0]) >= N
    tails = [subtrees[j if mode == orc.PARTIES else 0][-N:] for j in range(N)]          # :320-328, N clones without comm
    lt, start, layer = [], 0, N >> 1
    while layer > 0:                                                                     # :338-349
        for j in range(N):
            lt.append(tails[j][start:start + layer])
        start += layer
        layer >>= 1
    lt = list(np.concatenate(lt))
    for i in range(N * N - N, N * N - 1):                                                # :354-357
        x0, x1 = orc.sub_index(i)
        lt.append(orc.fr_mul(lt[x0].reshape(1, 4), lt[x1].reshape(1, 4))[0])
    lt.append(np.zeros(4, dtype=np.uint64))
    assert P == len(subtrees)
    return subtrees, np.array(lt, dtype=np.uint64)


def c_acc_product_and_share(pp, mode, shares, masks, unmask0, unmask1, unmask2):
    """dacc_product.rs:66-292.  Every argument is a list with one array per party (LEADER_SIM: one).
    -> list per party of (share0, share1, share2)"""
    N, l = pp.n, pp.l
    P = N if mode == orc.PARTIES else 1
    L = len(shares[0])
    assert L > N and L % N == 0
    block = L // N
    masked = [orc.fr_mul(shares[p], masks[p]) for p in range(P)]                         # :88-93
    # d_unpack2_many with receiver i (:95-107): party i ends with unpack2 of block i of everybody's masked shares
    masked_x = []
    for p in range(P):
        rows = [masked[j if mode == orc.PARTIES else 0][p * block:(p + 1) * block] for j in range(N)]
        masked_x.append(np.concatenate([orc.unpack2(pp, np.stack([rows[j][b] for j in range(N)])) for b in range(block)]))
    subtrees, ltree = c_acc_product(pp, mode, masked_x)                                  # :111-113
    m = block * l
    sm = []
    for p in range(P):
        st = subtrees[p]
        to_share = st[:len(st) - N]                                                      # :119
        sm.append((_pack_chunks(pp, to_share[0::2]), _pack_chunks(pp, to_share[1::2]),
                   _pack_chunks(pp, to_share[len(st) // 2:])))                           # :120-152
    out = []
    q0 = N * N // 2 // l
    lead = (_pack_chunks(pp, ltree[0::2]), _pack_chunks(pp, ltree[1::2]), _pack_chunks(pp, ltree))   # :212-249 (whole tree for v1x)
    for p in range(P):
        res = []
        for k in range(3):
            if mode == orc.PARTIES:
                rows = [sm[i][k][p] for i in range(N)]                                   # hub i sends row p (:154-193)
            else:
                rows = [sm[0][k][i] for i in range(N)]                                   # :196-202 placeholder
            merged = _merge(rows)                                                        # :204-209
            full = np.concatenate([merged, lead[k][p]])                                  # :251-262 (leader keeps row 0)
            assert len(full) == L, (len(merged), len(lead[k][p]), L, q0, m)
            res.append(orc.fr_mul(full, (unmask0, unmask1, unmask2)[k][p]))              # :265-275
        out.append(tuple(res))                                                           # degree_reduce_many results are dropped (:278-285)
    return out


def cpermcheck(n, pks, pp, mode, N, algo="ark"):
    """hyperplonk/src/dhyperplonk.rs:1249-1385.  pks: per party dict with V, sid, ssigma, eq_r1, mask, unmask0..2
    (4 * 2^n / l entries each), challenge_r1, alpha, beta, c_commitment.  -> per party dict like dhyperplonk's."""
    P = N if mode == orc.PARTIES else 1
    assert len(pks) == P
    csrs = [pk["c_commitment"] for pk in pks]
    res = [dict(gate_identity_proofs=[], gate_identity_commitments=[], wiring_proofs=[], wiring_commits=[],
                wiring_opens=[]) for _ in range(P)]
    col = lambda name: [pk[name] for pk in pks]   # noqa: E731
    r1 = pks[0]["challenge_r1"]

    def commit(tabs):
        o = orc.c_commit(csrs, pp, mode, [[t] for t in tabs], algo)
        for j in range(P):
            res[j]["wiring_commits"].append(o[j, 0:1].copy())

    def open_(tabs):
        val, proofs = orc.c_open(csrs, pp, mode, tabs, r1, algo)
        for j in range(P):
            res[j]["wiring_opens"].append((val[j:j + 1].copy(), proofs[j].copy()))

    def sumcheck(f, g):
        o = orc.c_sumcheck_product(pp, mode, f, g, r1)
        for j in range(P):
            res[j]["wiring_proofs"].append(o[j].copy())
    num, den = [], []
    for pk in pks:
        L = len(pk["V"])
        alpha, beta = _rep(pk["alpha"], L), _rep(pk["beta"], L)
        num.append(orc.fr_add(orc.fr_add(pk["V"], orc.fr_mul(alpha, pk["sid"])), beta))          # :1278-1280
        den.append(orc.fr_add(orc.fr_add(pk["eq_r1"], orc.fr_mul(alpha, pk["ssigma"])), beta))   # :1281-1283
    commit(col("ssigma"))
    open_(col("ssigma"))
    commit(col("sid"))
    open_(col("sid"))
    for ev in (num, den):                                                                        # :1311-1376
        sh = c_acc_product_and_share(pp, mode, ev, col("mask"), col("unmask0"), col("unmask1"), col("unmask2"))
        vx0, vx1, v1x = [s[0] for s in sh], [s[1] for s in sh], [s[2] for s in sh]
        commit(ev)
        open_(ev)
        commit(vx0)
        open_(vx0)
        commit(vx1)
        open_(vx1)
        commit(v1x)
        open_(v1x)
        sumcheck(col("eq_r1"), v1x)
        sumcheck(col("eq_r1"), vx0)
        sumcheck(vx0, vx1)
        open_(ev)
    return res


def local_hyperplonk(n, pk, algo="ark"):
    """hyperplonk/src/hyperplonk.rs:15-160 on explicit inputs.  pk: dict with m, a_evals, b_evals, c_evals, input, q1, q2,
    ssigma, sid, eq, eq_p2, challenge, challengep2, alpha, beta, commitment (oracle.Srs with levels 0 .. n+2)."""
    srs = pk["commitment"]
    res = dict(gate_identity_proofs=[], gate_identity_commitments=[], wiring_proofs=[], wiring_commits=[], wiring_opens=[])
    ch, ch2 = pk["challenge"], pk["challengep2"]
    gate_tabs = [pk[k] for k in ("a_evals", "b_evals", "c_evals", "input", "q1", "q2")]
    coms = [orc.commit(srs, t, algo) for t in gate_tabs]                                         # :55-62
    gp = res["gate_identity_proofs"]
    gp.append(orc.sumcheck_product(pk["eq"], pk["q1"], ch))                                      # :77
    gp.append(orc.sumcheck_product(pk["q1"], orc.fr_add(pk["a_evals"], pk["b_evals"]), ch))      # :78-84
    gp.append(orc.sumcheck_product(pk["eq"], pk["q2"], ch))
    gp.append(orc.sumcheck_product(pk["a_evals"], pk["b_evals"], ch))
    gp.append(orc.sumcheck_product(pk["q2"], pk["a_evals"], ch))
    gp.append(orc.sumcheck_product(pk["eq"], orc.fr_sub(pk["input"], pk["c_evals"]), ch))        # :90-96
    m = pk["m"]
    alpha, beta = _rep(pk["alpha"], len(m)), _rep(pk["beta"], len(m))
    num = orc.fr_add(orc.fr_add(m, orc.fr_mul(alpha, pk["sid"])), beta)                          # :106-110
    den = orc.fr_add(orc.fr_add(m, orc.fr_mul(alpha, pk["ssigma"])), beta)                       # :111-115
    h = orc.fr_mul(num, orc.fr_inv(den))                                                         # :116
    vx0, vx1, v1x = orc.acc_product(h)                                                           # :118
    for t in (pk["sid"], pk["ssigma"], h, num, den, vx0, vx1, v1x):                              # :121-136
        res["wiring_commits"].append(orc.commit(srs, t, algo))
        res["wiring_opens"].append(orc.open_(srs, t, ch2, algo))
    wp = res["wiring_proofs"]
    wp.append(orc.sumcheck_product(pk["eq_p2"], v1x, ch2))                                       # :138-145
    wp.append(orc.sumcheck_product(pk["eq_p2"], vx0, ch2))
    wp.append(orc.sumcheck_product(vx0, vx1, ch2))
    wp.append(orc.sumcheck_product(pk["eq_p2"], den, ch2))
    wp.append(orc.sumcheck_product(h, den, ch2))
    wp.append(orc.sumcheck_product(pk["eq_p2"], num, ch2))
    for com, t in zip(coms, gate_tabs):                                                          # :149-156
        res["gate_identity_commitments"].append((com, orc.open_(srs, t, ch)))
    return res
