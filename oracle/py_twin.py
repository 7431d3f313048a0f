"""ORACLE (test infrastructure) -- Python big-integer twin.

An independent, slow restatement used only to pin oracle/liboracle.so on small
cases: plain modular integers for Fr/Fq, affine chord-and-tangent for G1,
Lagrange evaluation for the PSS maps of secret-sharing/src/pss.rs.  Nothing in
the product imports it.
"""
R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
G1_X = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
G1_Y = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
GENERATOR = 7
TWO_ADICITY = 32
# known-answer constants recorded in BASELINE.md section 4 (l = 1, N = 8)
LAMBDA0 = 0x108faa3073a8c8c12be3b125b83bb125551b176e24920d246db6db6d92492493
MU0 = 0x3933589fbde041d8bc337099d1377b192b0b9e0f3afd3455bcf6ce47a15efcb9
OMEGA8 = 0x345766f603fa66e78c0625cd70d77ce2b38b21c28713b7007228fd3397743f7a

INF = None


def finv(a, m):
    return pow(a, m - 2, m)


# ------------------------------------------------------------------ G1 affine over python ints
def g1_add(p, q):
    if p is INF:
        return q
    if q is INF:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return INF
        lam = 3 * x1 * x1 * finv(2 * y1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * finv(x2 - x1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    return x3, (lam * (x1 - x3) - y1) % P_MOD


def g1_neg(p):
    return INF if p is INF else (p[0], (-p[1]) % P_MOD)


def g1_mul(p, k):
    acc, k = INF, k % R_MOD
    while k:
        if k & 1:
            acc = g1_add(acc, p)
        p = g1_add(p, p)
        k >>= 1
    return acc


def g1_msm(points, scalars):
    acc = INF
    for p, s in zip(points, scalars):
        acc = g1_add(acc, g1_mul(p, s))
    return acc


def on_curve(p):
    return p is INF or (p[1] * p[1] - p[0] ** 3 - 4) % P_MOD == 0


# ------------------------------------------------------------------ domains / PSS (pss.rs:38-171)
def root_of_unity(n):
    w = pow(GENERATOR, (R_MOD - 1) >> TWO_ADICITY, R_MOD)
    lg = n.bit_length() - 1
    assert 1 << lg == n
    return pow(w, 1 << (TWO_ADICITY - lg), R_MOD)


class Domain:
    def __init__(self, size, offset=1):
        self.size, self.offset, self.gen = size, offset % R_MOD, root_of_unity(size)

    def points(self):
        return [self.offset * pow(self.gen, j, R_MOD) % R_MOD for j in range(self.size)]


def _resize(v, n, zero):
    return list(v[:n]) + [zero] * (n - len(v))


class Ops:
    """element algebra: Fr ints or G1 affine points"""

    def __init__(self, kind):
        self.kind = kind
        self.zero = 0 if kind == "fr" else INF

    def add(self, a, b):
        return (a + b) % R_MOD if self.kind == "fr" else g1_add(a, b)

    def mul(self, a, k):
        return a * k % R_MOD if self.kind == "fr" else g1_mul(a, k)


def fft(dom, v, ops):
    v = _resize(v, dom.size, ops.zero)
    out = []
    for x in dom.points():
        acc, pw = ops.zero, 1
        for c in v:
            acc = ops.add(acc, ops.mul(c, pw))
            pw = pw * x % R_MOD
        out.append(acc)
    return out


def ifft(dom, v, ops):
    v = _resize(v, dom.size, ops.zero)
    n_inv, g_inv, o_inv = finv(dom.size, R_MOD), finv(dom.gen, R_MOD), finv(dom.offset, R_MOD)
    out = []
    for i in range(dom.size):
        acc = ops.zero
        for j, e in enumerate(v):
            acc = ops.add(acc, ops.mul(e, pow(g_inv, i * j, R_MOD)))
        out.append(ops.mul(acc, n_inv * pow(o_inv, i, R_MOD) % R_MOD))
    return out


class PSS:
    def __init__(self, l):
        self.l, self.n, self.t = l, 8 * l, l - 1
        self.share = Domain(self.n)
        self.secret = Domain(2 * l, GENERATOR)
        self.secret2 = Domain(4 * l, GENERATOR)

    def pack_from_public(self, secrets, kind="fr"):
        ops = Ops(kind)
        return fft(self.share, ifft(self.secret, secrets, ops), ops)

    def pack_single(self, s, kind="fr"):
        return self.pack_from_public(self.pack_from_public([s], kind), kind)

    def unpack(self, shares, kind="fr"):
        ops = Ops(kind)
        return fft(self.secret, ifft(self.share, shares, ops), ops)[: self.l]

    def unpack2(self, shares, kind="fr"):
        ops = Ops(kind)
        return fft(self.secret2, ifft(self.share, shares, ops), ops)[0:2 * self.l:2]


# ------------------------------------------------------------------ sumcheck (dsumcheck.rs:28-90)
def sumcheck_product(f, g, challenge):
    out, f, g = [], list(f), list(g)
    n = len(f).bit_length() - 1
    for i in range(n):
        h = len(f) // 2
        r = challenge[i]
        s0 = sum(f[j] * g[j] for j in range(h)) % R_MOD
        s1 = sum(f[h + j] * g[h + j] for j in range(h)) % R_MOD
        s2 = sum((2 * f[h + j] - f[j]) * (2 * g[h + j] - g[j]) for j in range(h)) % R_MOD
        out.append((s0, s1, s2))
        f = [(f[j] * (1 - r) + f[h + j] * r) % R_MOD for j in range(h)]
        g = [(g[j] * (1 - r) + g[h + j] * r) % R_MOD for j in range(h)]
    out.append((0, f[0] * g[0] % R_MOD, 0))
    return out


def check_sumcheck_product(h, proof, challenge, rounds):
    """the verifier identity of dsumcheck.rs:558-588"""
    if (proof[0][0] + proof[0][1]) % R_MOD != h % R_MOD:
        return False
    inv2 = finv(2, R_MOD)
    for i in range(1, rounds):
        p0, p1, p2 = proof[i - 1]
        x = challenge[i - 1]
        c = p0
        b = (-p2 + 4 * p1 - 3 * p0) * inv2 % R_MOD
        a = (p2 - 2 * p1 + p0) * inv2 % R_MOD
        if (proof[i][0] + proof[i][1]) % R_MOD != (a * x * x + b * x + c) % R_MOD:
            return False
    return True


def mle_eval(evals, point):
    """fix variables top-down like mle.rs:88-104"""
    v = list(evals)
    for u in point:
        h = len(v) // 2
        v = [(v[j] * (1 - u) + v[h + j] * u) % R_MOD for j in range(h)]
    return v[0]


def sub_index(i):
    first = i.bit_length() - 1
    x = (i & ~(1 << first)) << 1
    return x, x + 1


def acc_product(x):
    t = list(x) + list(x)
    for i in range(len(x), 2 * len(x) - 1):
        a, b = sub_index(i)
        t[i] = t[a] * t[b] % R_MOD
    t[2 * len(x) - 1] = 0
    return t[0::2], t[1::2], t[len(t) // 2:]


# ---------------------------------------------------------------- wire format (ark-bls12-381 0.4.0 compressed G1 = Zcash / IETF)
def g1_serialize_compressed(pt):
    """pt: (x, y) ints or None (infinity) -> 48 bytes: x big endian, flags compressed | infinity | y larger root"""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if y > (P_MOD - y) % P_MOD else 0)
    return bytes(b)


def g1_deserialize_compressed(b):
    """48 bytes -> ((x, y) or None, status): 0 ok, 1 malformed / not on the curve, 2 not in the r-torsion"""
    assert len(b) == 48
    flags = b[0] & 0xE0
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if not flags & 0x80:
        return None, 1
    if flags & 0x40:
        return (None, 0) if (not flags & 0x20 and x == 0) else (None, 1)
    if x >= P_MOD:
        return None, 1
    rhs = (x * x * x + 4) % P_MOD
    y = pow(rhs, (P_MOD + 1) // 4, P_MOD)
    if y * y % P_MOD != rhs:
        return None, 1
    if (y > (P_MOD - y) % P_MOD) != bool(flags & 0x20):
        y = (P_MOD - y) % P_MOD
    acc, q, k = INF, (x, y), R_MOD          # [r]P without reducing the scalar (g1_mul works mod r)
    while k:
        if k & 1:
            acc = g1_add(acc, q)
        q = g1_add(q, q)
        k >>= 1
    if acc is not INF:
        return None, 2
    return (x, y), 0


# ---------------------------------------------------------------- G2 (for d_msm over G2; Fq2 = Fq[u] / (u^2 + 1))
# public constants of BLS12-381: the G2 generator, curve y^2 = x^3 + 4 (1 + u)
G2_X = (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
        0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e)
G2_Y = (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
        0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)
G2_B = (4, 4)


def f2_add(a, b): return ((a[0] + b[0]) % P_MOD, (a[1] + b[1]) % P_MOD)
def f2_sub(a, b): return ((a[0] - b[0]) % P_MOD, (a[1] - b[1]) % P_MOD)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P_MOD, (a[0] * b[1] + a[1] * b[0]) % P_MOD)


def f2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P_MOD, P_MOD - 2, P_MOD)
    return (a[0] * n % P_MOD, -a[1] * n % P_MOD)


def g2_on_curve(p):
    if p is None:
        return True
    x, y = p
    return f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), G2_B)


def g2_add(p, q):
    """affine addition over Fq2; None is the identity"""
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if f2_add(p[1], q[1]) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(p[0], p[0])), f2_inv(f2_add(p[1], p[1])))
    else:
        lam = f2_mul(f2_sub(q[1], p[1]), f2_inv(f2_sub(q[0], p[0])))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), p[0]), q[0])
    return (x3, f2_sub(f2_mul(lam, f2_sub(p[0], x3)), p[1]))


def g2_mul(p, k):
    acc = None
    for bit in bin(k % R_MOD)[2:] if k % R_MOD else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, p)
    return acc
