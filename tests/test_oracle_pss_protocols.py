"""Pins the oracle's PSS maps and protocols: the reference's own property tests
(secret-sharing/src/pss.rs:191-288, dist-primitive/src/dmsm.rs:72-138,
dsumcheck.rs:541-588,687-747, dacc_product.rs:442-466, dpoly_comm.rs:511-531)
re-expressed, the closed forms of BASELINE.md section 4, and the big-int twin.  CPU only."""
import numpy as np
import pytest

from oracle import py_twin as tw


def _rand_ints(rng, n, mod=tw.R_MOD):
    return [int.from_bytes(rng.bytes(64), "little") % mod for _ in range(n)]


def _g1c(p):
    return (0, 0, 1) if p is None else (p[0], p[1], 0)


# ---------------------------------------------------------------- PSS
def test_initialize(orc):                      # pss.rs:191-200
    pp = orc.pp_new(4)
    assert (pp.t, pp.l, pp.n) == (3, 4, 32)
    assert orc.fr_to_ints(np.array([pp.share_gen[:]], dtype=np.uint64)) == [tw.root_of_unity(32)]
    pp1 = orc.pp_new(1)
    assert orc.fr_to_ints(np.array([pp1.share_gen[:]], dtype=np.uint64)) == [tw.OMEGA8]


@pytest.mark.parametrize("l", [1, 2, 4])
def test_pack_unpack_fr_vs_twin(orc, l):       # pss.rs:202-234
    rng = np.random.default_rng(10 + l)
    pp, ps = orc.pp_new(l), tw.PSS(l)
    sec = _rand_ints(rng, l)
    sh = orc.pack_from_public(pp, orc.fr_from_ints(sec))
    assert orc.fr_to_ints(sh) == ps.pack_from_public(sec)
    assert orc.fr_to_ints(orc.unpack(pp, sh)) == sec
    sec2 = _rand_ints(rng, l)
    sh2 = orc.pack_from_public(pp, orc.fr_from_ints(sec2))
    prod = orc.fr_mul(sh, sh2)
    assert orc.fr_to_ints(orc.unpack2(pp, prod)) == [a * b % tw.R_MOD for a, b in zip(sec, sec2)]
    assert orc.fr_to_ints(orc.unpack2(pp, prod)) == ps.unpack2(orc.fr_to_ints(prod))
    # pack_single (pss.rs:103-113): the double packing
    assert orc.fr_to_ints(orc.pack_single(pp, orc.fr_from_ints(sec[:1]))) == ps.pack_single(sec[0])
    # unpack of arbitrary (non-codeword) shares follows the truncation semantics
    junk = _rand_ints(rng, 8 * l)
    assert orc.fr_to_ints(orc.unpack(pp, orc.fr_from_ints(junk))) == ps.unpack(junk)
    assert orc.fr_to_ints(orc.unpack2(pp, orc.fr_from_ints(junk))) == ps.unpack2(junk)


def test_closed_forms_l1(orc):
    pp = orc.pp_new(1)
    s = 0x1234567890abcdef1234567890abcdef
    sh = orc.fr_to_ints(orc.pack_from_public(pp, orc.fr_from_ints([s])))
    inv2, inv7 = pow(2, -1, tw.R_MOD), pow(7, -1, tw.R_MOD)
    lam = [(1 + pow(tw.OMEGA8, j, tw.R_MOD) * inv7) * inv2 % tw.R_MOD for j in range(8)]
    assert lam[0] == 4 * inv7 % tw.R_MOD == tw.LAMBDA0
    assert sh == [s * x % tw.R_MOD for x in lam]
    one = orc.fr_to_ints(orc.pack_single(pp, orc.fr_from_ints([1])))
    assert one[0] == tw.MU0
    # leader-sim closed forms
    x = orc.fr_from_ints([s])
    assert orc.fr_to_ints(orc.pss2ss(pp, orc.LEADER_SIM, x)[0]) == [tw.MU0 * s % tw.R_MOD]
    assert orc.fr_to_ints(orc.degree_reduce(pp, orc.LEADER_SIM, x)) == [tw.LAMBDA0 * s % tw.R_MOD]


def test_group_addition_g1(orc):               # pss.rs:236-254 over G1
    rng = np.random.default_rng(20)
    l = 2
    pp, ps = orc.pp_new(l), tw.PSS(l)
    G = (tw.G1_X, tw.G1_Y)
    ks = _rand_ints(rng, l)
    sec_pts = [tw.g1_mul(G, k) for k in ks]
    sec = orc.g1_from_affine(orc.g1_gen_mul(orc.fr_from_ints(ks)))
    sh = orc.pack_from_public(pp, sec, kind=1)
    assert orc.canon_g1(sh) == [_g1c(p) for p in ps.pack_from_public(sec_pts, "g1")]
    assert orc.canon_g1(orc.unpack(pp, sh, kind=1)) == [_g1c(p) for p in sec_pts]
    dbl = orc.g1_add(sh, sh)
    assert orc.canon_g1(orc.unpack2(pp, dbl, kind=1)) == [_g1c(tw.g1_add(p, p)) for p in sec_pts]


# ---------------------------------------------------------------- d_msm
def test_pack_unpack2_msm_property(orc):       # dmsm.rs:92-138 with L=2, N=16, smaller M
    rng = np.random.default_rng(30)
    l, M = 2, 16
    pp = orc.pp_new(l)
    G = (tw.G1_X, tw.G1_Y)
    ks, ss = _rand_ints(rng, M), _rand_ints(rng, M)
    want = _g1c(tw.g1_mul(G, sum(k * s for k, s in zip(ks, ss)) % tw.R_MOD))
    gsec = orc.g1_from_affine(orc.g1_gen_mul(orc.fr_from_ints(ks)))
    fsec = orc.fr_from_ints(ss)
    gsh = np.stack([orc.pack_from_public(pp, gsec[i:i + l], kind=1) for i in range(0, M, l)])   # (M/l, N, 18)
    fsh = np.stack([orc.pack_from_public(pp, fsec[i:i + l]) for i in range(0, M, l)])           # (M/l, N, 4)
    bases = [[orc.g1_to_affine(gsh[:, p])] for p in range(pp.n)]
    scal = [[np.ascontiguousarray(fsh[:, p])] for p in range(pp.n)]
    # the reference's property: sum(unpack2(per-party msm)) == msm(plain)
    per_party = np.concatenate([orc.msm(bases[p][0], scal[p][0]) for p in range(pp.n)])
    got = orc.unpack2(pp, per_party, kind=1)
    acc = got[:1]
    for i in range(1, l):
        acc = orc.g1_add(acc, got[i:i + 1])
    assert orc.canon_g1(acc) == [want]
    # d_msm in PARTIES mode: each party's output is a packed share of [sum]*l -> unpack gives sum twice
    out = orc.d_msm(pp, orc.PARTIES, bases, scal)                    # (N, 1, 18)
    un = orc.unpack(pp, np.ascontiguousarray(out[:, 0]), kind=1)
    assert orc.canon_g1(un) == [want] * l
    assert orc.canon_g1(orc.d_msm(pp, orc.PARTIES, bases, scal, algo="naive")[:, 0]) == orc.canon_g1(out[:, 0])


def test_d_msm_leader_sim_closed_form(orc):    # out = lambda0 * msm
    rng = np.random.default_rng(31)
    pp = orc.pp_new(1)
    G = (tw.G1_X, tw.G1_Y)
    for m in (1, 40):
        ks, ss = _rand_ints(rng, m), _rand_ints(rng, m)
        bases = orc.g1_gen_mul(orc.fr_from_ints(ks))
        out = orc.d_msm(pp, orc.LEADER_SIM, [[bases]], [[orc.fr_from_ints(ss)]])
        dot = sum(k * s for k, s in zip(ks, ss)) % tw.R_MOD
        assert orc.canon_g1(out[0]) == [_g1c(tw.g1_mul(G, dot * tw.LAMBDA0 % tw.R_MOD))]


# ---------------------------------------------------------------- sumcheck
def test_sumcheck_product_vs_twin_and_verifier(orc):     # dsumcheck.rs:687-747
    rng = np.random.default_rng(40)
    n = 6
    f, g, ch = _rand_ints(rng, 1 << n), _rand_ints(rng, 1 << n), _rand_ints(rng, n)
    got = orc.sumcheck_product(orc.fr_from_ints(f), orc.fr_from_ints(g), orc.fr_from_ints(ch))
    got = [tuple(orc.fr_to_ints(t)) for t in got]
    assert got == tw.sumcheck_product(f, g, ch)
    h = sum(a * b for a, b in zip(f, g)) % tw.R_MOD
    assert tw.check_sumcheck_product(h, got, ch, n)
    assert got[-1][1] == tw.mle_eval(f, ch) * tw.mle_eval(g, ch) % tw.R_MOD


def test_c_sumcheck_product_parties(orc):      # dsumcheck.rs:809-862 re-expressed with N = 8l parties
    rng = np.random.default_rng(41)
    l, n = 2, 4
    pp = orc.pp_new(l)
    f, g, ch = _rand_ints(rng, l << n), _rand_ints(rng, l << n), _rand_ints(rng, n + 1)
    F, Gg = orc.fr_from_ints(f), orc.fr_from_ints(g)
    # packing layout of examples/sumcheck.rs: share vector position i packs secrets {i + k*2^n}, k<l
    fsh = np.stack([orc.pack_from_public(pp, F[i::1 << n]) for i in range(1 << n)])     # (2^n, N, 4)
    gsh = np.stack([orc.pack_from_public(pp, Gg[i::1 << n]) for i in range(1 << n)])
    fs = [np.ascontiguousarray(fsh[:, p]) for p in range(pp.n)]
    gs = [np.ascontiguousarray(gsh[:, p]) for p in range(pp.n)]
    out = orc.c_sumcheck_product(pp, orc.PARTIES, fs, gs, orc.fr_from_ints(ch))          # (N, n+2, 3, 4)
    # Phase-1 triples are degree-2(t+l) sharings: unpack2 across parties and sum the l secrets
    for i in range(n):
        for c in range(3):
            col = np.ascontiguousarray(out[:, i, c])
            tot = sum(orc.fr_to_ints(orc.unpack2(pp, col))) % tw.R_MOD
            if i == 0 and c == 0:
                first0 = tot
            if i == 0 and c == 1:
                first1 = tot
    assert (first0 + first1) % tw.R_MOD == sum(a * b for a, b in zip(f, g)) % tw.R_MOD
    # leader-sim is deterministic and equals party 0's local phase-1 view
    ls = orc.c_sumcheck_product(pp, orc.LEADER_SIM, fs[:1], gs[:1], orc.fr_from_ints(ch))
    assert np.array_equal(ls[0, :n], out[0, :n])


def test_d_sumcheck_product(orc):              # dsumcheck.rs:359-512: slices + leader tail == monolithic
    rng = np.random.default_rng(42)
    N, n = 8, 4
    tot = N << n
    f, g, ch = _rand_ints(rng, tot), _rand_ints(rng, tot), _rand_ints(rng, n + 3)
    F, Gg = orc.fr_from_ints(f), orc.fr_from_ints(g)
    fs = [F[p << n:(p + 1) << n] for p in range(N)]
    gs = [Gg[p << n:(p + 1) << n] for p in range(N)]
    out = orc.d_sumcheck_product(orc.PARTIES, N, fs, gs, orc.fr_from_ints(ch))
    got = [tuple(orc.fr_to_ints(t)) for t in out]
    assert len(got) == n + 3
    # the distributed protocol binds the LOCAL variables first, then the party index
    chal_mono = ch[n:] + ch[:n]
    # monolithic check through the verifier identity with h = <f,g>
    h = sum(a * b for a, b in zip(f, g)) % tw.R_MOD
    assert (got[0][0] + got[0][1]) % tw.R_MOD == h
    assert tw.check_sumcheck_product(h, got, ch, n + 3)
    del chal_mono


# ---------------------------------------------------------------- product tree
def test_sub_index_and_acc_product_known_answers(orc):   # dacc_product.rs:442-466
    assert orc.sub_index(26) == (20, 21) == tw.sub_index(26)
    v0, v1, v2 = orc.acc_product(orc.fr_from_ints([1, 2, 3, 4]))
    assert orc.fr_to_ints(v0) == [1, 3, 2, 24]
    assert orc.fr_to_ints(v1) == [2, 4, 12, 0]
    assert orc.fr_to_ints(v2) == [2, 12, 24, 0]
    rng = np.random.default_rng(50)
    x = _rand_ints(rng, 32)
    got = orc.acc_product(orc.fr_from_ints(x))
    assert tuple(orc.fr_to_ints(v) for v in got) == tuple(tw.acc_product(x))


def test_d_acc_product(orc):
    rng = np.random.default_rng(51)
    N, m = 8, 16
    xs = [orc.fr_from_ints(_rand_ints(rng, m)) for _ in range(N)]
    subs, top = orc.d_acc_product(orc.PARTIES, N, xs)
    for p in range(N):
        assert np.array_equal(subs[p], orc.acc_product_tree(xs[p]))
    # the reference sends subtree[last] AFTER forcing it to zero (dacc_product.rs:381,390)
    assert orc.fr_to_ints(top) == [0] * (2 * N)


# ---------------------------------------------------------------- PST commit / open
def test_commit_open_trapdoor(orc):            # dpoly_comm.rs:511-531 with the pairing replaced by the trapdoor
    rng = np.random.default_rng(60)
    n = 5
    s, pe, u = _rand_ints(rng, n), _rand_ints(rng, 1 << n), _rand_ints(rng, n)
    G = (tw.G1_X, tw.G1_Y)
    srs = orc.Srs.new(orc.g1_from_affine(orc.g1_generator()), orc.fr_from_ints(s))
    assert srs.levels == n + 1
    # level n holds eq(s, x)*G ordered so that s_0 is the outermost (top) variable
    com = orc.commit(srs, orc.fr_from_ints(pe))
    assert orc.canon_g1(com) == [_g1c(tw.g1_mul(G, tw.mle_eval(pe, s)))]
    val, proofs = orc.open_(srs, orc.fr_from_ints(pe), orc.fr_from_ints(u))
    assert orc.fr_to_ints(val) == [tw.mle_eval(pe, u)]
    # verify() in the exponent: p(s) - v = sum_i q_i(s_{i+1..}) * (s_i - u_i)
    cur, acc = list(pe), 0
    for i in range(n):
        h = len(cur) // 2
        q = [(cur[h + j] - cur[j]) % tw.R_MOD for j in range(h)]
        qs = tw.mle_eval(q, s[i + 1:]) if h > 1 else q[0]
        assert orc.canon_g1(proofs[i:i + 1]) == [_g1c(tw.g1_mul(G, qs))]
        acc = (acc + qs * (s[i] - u[i])) % tw.R_MOD
        cur = [(cur[j] * (1 - u[i]) + cur[h + j] * u[i]) % tw.R_MOD for j in range(h)]
    assert (tw.mle_eval(pe, s) - tw.mle_eval(pe, u)) % tw.R_MOD == acc


def test_d_commit_d_open_equal_monolithic(orc):          # intent of dpoly_comm.rs:533-583 (with correct slices)
    rng = np.random.default_rng(61)
    N, n = 8, 3
    tot_vars = n + 3
    s, pe, u = _rand_ints(rng, tot_vars), _rand_ints(rng, 1 << tot_vars), _rand_ints(rng, tot_vars)
    g = orc.g1_from_affine(orc.g1_generator())
    full = orc.Srs.new(g, orc.fr_from_ints(s))
    PE = orc.fr_from_ints(pe)
    # party p holds slice p and the SRS slice for it: bases = top level restricted to its block;
    # lower levels of a per-party SRS are the trapdoor SRS over the local variables scaled by eq(s_top, p)
    want_com = orc.commit(full, PE)
    want_val, want_proofs = orc.open_(full, PE, orc.fr_from_ints(u))
    srs_list = []
    for p in range(N):
        bits = [(p >> (2 - k)) & 1 for k in range(3)]
        w = 1
        for k in range(3):
            w = w * (s[k] if bits[k] else (1 - s[k])) % tw.R_MOD
        gp = orc.g1_mul(g, orc.fr_from_ints([w]))
        loc = orc.Srs.new(gp, orc.fr_from_ints(s[3:]))
        srs_list.append(loc)
    # party 0's SRS is also used for the leader's root open over N values (dpoly_comm.rs:377): needs level 3 of a
    # trapdoor SRS over s[:3]; so only commitments and the LOCAL part of the proofs are comparable here
    slices = [PE[p << n:(p + 1) << n] for p in range(N)]
    com = orc.d_commit(srs_list, orc.PARTIES, N, slices)
    assert orc.canon_g1(com) == orc.canon_g1(want_com)
    val, proofs = orc.d_open(srs_list, orc.PARTIES, N, slices, orc.fr_from_ints(u))
    assert orc.fr_to_ints(val) == orc.fr_to_ints(want_val)
    assert len(proofs) == 3 + n
    # The summed local proofs are quotients of the LOCAL variables taken before the top variables are bound, so
    # they differ from the monolithic ones but satisfy the same verification equation restricted to the local
    # variables:  p(s) - p(s_top, u_loc) = sum_{i>=3} (s_i - u_i) * pi_i   (checked in the exponent via G1 ops)
    G = (tw.G1_X, tw.G1_Y)
    z = [tw.mle_eval(pe[p << n:(p + 1) << n], u[3:]) for p in range(N)]
    lhs = (tw.mle_eval(pe, s) - tw.mle_eval(z, s[:3])) % tw.R_MOD
    acc = None
    pis = orc.canon_g1(proofs[3:])
    for i in range(n):
        acc = tw.g1_add(acc, tw.g1_mul((pis[i][0], pis[i][1]), (s[3 + i] - u[3 + i]) % tw.R_MOD))
    assert _g1c(acc) == _g1c(tw.g1_mul(G, lhs))
    del want_proofs


def test_c_commit_c_open_leader_sim(orc):
    rng = np.random.default_rng(62)
    pp = orc.pp_new(1)
    n = 4
    s, pe, u = _rand_ints(rng, n), _rand_ints(rng, 1 << n), _rand_ints(rng, n)
    G = (tw.G1_X, tw.G1_Y)
    srs = orc.Srs.new(orc.g1_from_affine(orc.g1_generator()), orc.fr_from_ints(s))
    com = orc.c_commit([srs], pp, orc.LEADER_SIM, [[orc.fr_from_ints(pe)]])
    assert orc.canon_g1(com[0]) == [_g1c(tw.g1_mul(G, tw.mle_eval(pe, s) * tw.LAMBDA0 % tw.R_MOD))]
    val, proofs = orc.c_open([srs], pp, orc.LEADER_SIM, [orc.fr_from_ints(pe)], orc.fr_from_ints(u))
    assert orc.fr_to_ints(val) == [tw.mle_eval(pe, u) * tw.MU0 % tw.R_MOD]
    v2, p2 = orc.open_(srs, orc.fr_from_ints(pe), orc.fr_from_ints(u))
    lam = orc.fr_from_ints([tw.LAMBDA0])
    assert orc.canon_g1(proofs[0]) == orc.canon_g1(orc.g1_mul(p2, np.repeat(lam, n, axis=0)))


def test_fix_variable(orc):                    # mle.rs:88-104
    rng = np.random.default_rng(70)
    e, pts = _rand_ints(rng, 32), _rand_ints(rng, 2)
    got = orc.fr_to_ints(orc.fix_variable(orc.fr_from_ints(e), orc.fr_from_ints(pts)))
    v = list(e)
    for u in pts:
        h = len(v) // 2
        v = [(v[j] * (1 - u) + v[h + j] * u) % tw.R_MOD for j in range(h)]
    assert got == v
    a = orc.fr_to_ints(orc.fix_variable(orc.fr_from_ints(e), orc.fr_from_ints([0, 1])))
    assert a == e[8:16]


def test_single_mle_sumcheck_properties(orc):
    """sumcheck (dsumcheck.rs:6-26): round i's two sums add up to the previous round's polynomial at the challenge,
    round 0 sums to the table total, the final pair is (0, f(challenge)) = (0, fix_variable(f, challenge));
    d_sumcheck over 8 slices of a table equals sumcheck of the whole table with the slice index as the top variables
    (the reference's dsumcheck_test, :617-645, re-expressed)"""
    import numpy as np
    from oracle import py_twin as tw
    R = tw.R_MOD
    rng = np.random.default_rng(77)
    nv = 7
    f = orc.random_fr(rng, 1 << nv)
    ch = orc.random_fr(rng, nv)
    out = orc.sumcheck(f, ch)
    fi, ci = orc.fr_to_ints(f), orc.fr_to_ints(ch)
    pairs = [[orc.fr_to_ints(out[i, j:j + 1])[0] for j in range(2)] for i in range(nv + 1)]
    assert (pairs[0][0] + pairs[0][1]) % R == sum(fi) % R
    for i in range(nv):
        a, b = pairs[i]
        val = (a * (1 - ci[i]) + b * ci[i]) % R
        nxt = (pairs[i + 1][0] + pairs[i + 1][1]) % R
        assert val == nxt, i
    assert pairs[nv][0] == 0 and pairs[nv][1] == orc.fr_to_ints(orc.fix_variable(f, ch))[0]
    # distributed: party j holds slice j (the top three variables select the party)
    N, s = 8, 3
    slices = [f[j * (len(f) // N):(j + 1) * (len(f) // N)] for j in range(N)]
    ch_d = np.concatenate([ch[s:], ch[:s]])          # local rounds use challenge[..n], the leader challenge[n..n+s]
    lead = orc.d_sumcheck(orc.PARTIES, N, slices, ch_d)
    assert len(lead) == nv
    # the first local round's sums over all parties = sums over (x_top3 free, x_3 = 0 / 1)
    lo = sum(sum(orc.fr_to_ints(sl[:len(sl) // 2])) for sl in slices) % R
    hi = sum(sum(orc.fr_to_ints(sl[len(sl) // 2:])) for sl in slices) % R
    assert orc.fr_to_ints(lead[0, 0:1])[0] == lo and orc.fr_to_ints(lead[0, 1:2])[0] == hi
    # last leader round folds to the full evaluation: a*(1-r) + b*r at the final challenge = f(point)
    a, b = orc.fr_to_ints(lead[nv - 1, 0:1])[0], orc.fr_to_ints(lead[nv - 1, 1:2])[0]
    r_last = orc.fr_to_ints(ch_d[nv - 1:nv])[0]
    point = np.concatenate([ch_d[nv - s:], ch_d[:nv - s]])   # top variables first for fix_variable
    assert (a * (1 - r_last) + b * r_last) % R == orc.fr_to_ints(orc.fix_variable(f, point))[0]
