"""CPU ORACLE (test infrastructure, NOT product code) -- pure-Python restatement.

PARITY UNPINNED: the reference (gio54321/quill-zkvm) is Rust on arkworks 0.5.0 and
cannot be compiled in this image (no cargo/rustc), and its tests hold no byte-level
golden vectors.  This file restates the reference's algorithms with Python ints and
the `blake3` wheel (which wraps the same upstream Rust crate the reference uses,
blake3 1.8.2 / Cargo.lock:167).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it.

Each function cites the reference file:line it follows (paths relative to the
reference root).  arkworks-0.5 behaviours that are recalled, not readable here, are
isolated in the `ser_*` functions and `fr_from_le_bytes_mod_order`.
"""
from __future__ import annotations

import blake3 as _blake3

# ---------------------------------------------------------------- BN254 constants
FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
B_COEFF = 3  # y^2 = x^3 + 3
G1_GEN = (1, 2)
R256 = 1 << 256


# ---------------------------------------------------------------- serialization (ark-serialize 0.5.0)
def ser_usize(v: int) -> bytes:
    """usize -> u64 little-endian (ark-serialize impl for usize)."""
    return int(v).to_bytes(8, "little")


def ser_fr(v: int) -> bytes:
    """Fr uncompressed = 32 B little-endian canonical (non-Montgomery)."""
    return (v % FR).to_bytes(32, "little")


def ser_fr_vec(vs) -> bytes:
    """Vec<Fr> / &[Fr] / DensePolynomial{coeffs} = u64 LE length ‖ elements."""
    return ser_usize(len(vs)) + b"".join(ser_fr(v) for v in vs)


def ser_g1(p) -> bytes:
    """G1Projective/G1Affine uncompressed: x ‖ y (32 B LE each), flags OR-ed into the
    last byte: 0x80 if y > -y (i.e. y > (q-1)/2), 0x40 and zero coordinates at infinity."""
    if p is None:
        out = bytearray(64)
        out[63] |= 0x40
        return bytes(out)
    x, y = p
    out = bytearray(x.to_bytes(32, "little") + y.to_bytes(32, "little"))
    if y > (FQ - y) % FQ:
        out[63] |= 0x80
    return bytes(out)


def fr_from_le_bytes_mod_order(b: bytes) -> int:
    """PrimeField::from_le_bytes_mod_order."""
    return int.from_bytes(b, "little") % FR


def trim(coeffs):
    """DensePolynomial invariant: no trailing zero coefficients (zero poly = [])."""
    c = [x % FR for x in coeffs]
    while c and c[-1] == 0:
        c.pop()
    return c


# ---------------------------------------------------------------- transcript
class Transcript:
    """transcript/src/transcript.rs:14-75."""

    def __init__(self, domain: bytes):
        self.domain = bytes(domain)
        self.state = _blake3.blake3(self.domain).digest()  # :15-23

    def append_bytes(self, message: bytes):  # :26-32
        self.state = _blake3.blake3(self.state + bytes(message)).digest()

    def append_usize(self, v):
        self.append_bytes(ser_usize(v))

    def append_fr(self, v):
        self.append_bytes(ser_fr(v))

    def append_fr_vec(self, vs):
        self.append_bytes(ser_fr_vec(vs))

    def append_g1(self, p):
        self.append_bytes(ser_g1(p))

    def draw_challenge(self, n: int) -> bytes:  # :49-63
        c = _blake3.blake3(self.state + b"challenge").digest(length=n)
        self.append_bytes(c)
        return c

    def draw_field_element(self) -> int:  # :71-75
        num_bytes = (254 + 128 + 7) // 8
        return fr_from_le_bytes_mod_order(self.draw_challenge(num_bytes))


# ---------------------------------------------------------------- univariate polys over Fr
def poly_add(a, b):
    n = max(len(a), len(b))
    return trim([(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)])


def poly_mul(a, b):
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % FR
    return trim(out)


def poly_eval(a, x):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % FR
    return acc


# ---------------------------------------------------------------- expression tree
# hyperplonk/src/utils/virtual_polynomial.rs:9-18.  Tuples: ("in", i) ("const", c) ("add", l, r) ("mul", l, r)
def e_in(i):
    return ("in", i)


def e_const(c):
    return ("const", c % FR)


def e_add(l, r):
    return ("add", l, r)


def e_mul(l, r):
    return ("mul", l, r)


def e_sub(l, r):
    """virtual_polynomial.rs:67-77 / 243-256: l + (-1)*r."""
    return ("add", l, ("mul", ("const", FR - 1), r))


def expr_eval_point(e, g):  # virtual_polynomial.rs:22-37
    t = e[0]
    if t == "in":
        return g[e[1]] % FR
    if t == "const":
        return e[1]
    l, r = expr_eval_point(e[1], g), expr_eval_point(e[2], g)
    return (l + r) % FR if t == "add" else (l * r) % FR


def expr_eval_poly(e, gp):  # virtual_polynomial.rs:300-320
    t = e[0]
    if t == "in":
        return list(gp[e[1]])
    if t == "const":
        return trim([e[1]])
    l, r = expr_eval_poly(e[1], gp), expr_eval_poly(e[2], gp)
    return poly_add(l, r) if t == "add" else poly_mul(l, r)


# ---------------------------------------------------------------- sumcheck
def sumcheck_prove(num_vars, tables, h, claimed_sum, tr: Transcript):
    """hyperplonk/src/piops/sumcheck.rs:28-114.  Returns (r_polys, point, evaluation)."""
    tr.append_usize(num_vars)  # :35
    tr.append_fr(claimed_sum)  # :36
    gs = [list(t) for t in tables]  # :44-49
    r_polys, point, evaluation = [], [], 0
    for i in reversed(range(num_vars)):  # :51
        msg = []
        lin = []
        for p in range(1 << i):  # :53-63
            polys = [trim([g[2 * p], g[2 * p + 1] - g[2 * p]]) for g in gs]
            lin.append(polys)
            msg = poly_add(msg, expr_eval_poly(h, polys))  # :67-70
        tr.append_fr_vec(msg)  # :73
        r_polys.append(msg)
        r = tr.draw_field_element()  # :77
        point.append(r)
        gs = [[poly_eval(lin[p][k], r) for p in range(1 << i)] for k in range(len(gs))]  # :81-92
        if i == 0:  # :94-100
            evaluation = expr_eval_point(h, [g[0] for g in gs])
    if num_vars == 0:
        pass  # reference leaves evaluation_claim = 0 when there are no rounds (:41)
    return r_polys, point, evaluation


def sumcheck_verify(num_vars, claimed_sum, r_polys, tr: Transcript):
    """sumcheck.rs:116-150.  Returns (point, evaluation) or raises ValueError."""
    tr.append_usize(num_vars)
    tr.append_fr(claimed_sum)
    v = claimed_sum % FR
    point = []
    for poly in r_polys:
        if (poly_eval(poly, 0) + poly_eval(poly, 1)) % FR != v:
            raise ValueError("Sumcheck polynomial does not sum to previous value")
        tr.append_fr_vec(poly)
        r = tr.draw_field_element()
        point.append(r)
        v = poly_eval(poly, r)
    return point, v


# ---------------------------------------------------------------- eq table / zero-check
def fast_eq_eval_hypercube(n, point):  # hyperplonk/src/utils/eq_eval.rs:6-31
    assert len(point) == n
    evals = [1]
    for i in reversed(range(n)):
        r = point[i] % FR
        om = (1 - r) % FR
        new = []
        for e in evals:
            new.append(e * om % FR)
            new.append(e * r % FR)
        evals = new
    return evals


def eq_eval(x, r):  # eq_eval.rs:33-43
    res = 1
    for a, b in zip(x, r):
        res = res * ((a * b + (1 - a) * (1 - b)) % FR) % FR
    return res


def fr_inv(a):
    return pow(a % FR, FR - 2, FR)


def zerocheck_prove(num_vars, tables, h, tr: Transcript):
    """hyperplonk/src/piops/zerocheck.rs:14-49.  Returns (r_polys, point, evaluation, z)."""
    z = [tr.draw_field_element() for _ in range(num_vars)]  # :20-22
    eq = fast_eq_eval_hypercube(num_vars, z)  # :25
    tables2 = list(tables) + [eq]  # :27
    h_hat = e_mul(h, e_in(len(tables)))  # :28-29
    r_polys, point, ev = sumcheck_prove(num_vars, tables2, h_hat, 0, tr)  # :31-32
    e = eq_eval(z, point)  # :34
    return r_polys, point, ev * fr_inv(e) % FR, z  # :36-40


def mle_evaluate(evals, point):
    """ark-poly DenseMultilinearExtension::evaluate: variable j <-> index bit j."""
    t = [e % FR for e in evals]
    for r in point:
        t = [(t[2 * p] + r * (t[2 * p + 1] - t[2 * p])) % FR for p in range(len(t) // 2)]
    assert len(t) == 1
    return t[0]


# ---------------------------------------------------------------- G1 (affine, Python ints; None = infinity)
def g1_is_on_curve(p):
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - B_COEFF) % FQ == 0


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % FQ)


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % FQ == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, FQ - 2, FQ) % FQ
    else:
        lam = (y2 - y1) * pow(x2 - x1, FQ - 2, FQ) % FQ
    x3 = (lam * lam - x1 - x2) % FQ
    return (x3, (lam * (x1 - x3) - y1) % FQ)


def g1_mul(p, k):
    k %= FR
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, p)
        p = g1_add(p, p)
        k >>= 1
    return acc


def msm_naive(bases, scalars):
    """VariableBaseMSM::msm_unchecked semantics (called at pcs/src/kzg.rs:72): zip to the shorter."""
    acc = None
    for b, s in zip(bases, scalars):
        acc = g1_add(acc, g1_mul(b, s))
    return acc


# ---------------------------------------------------------------- KZG
class KZG:
    """pcs/src/kzg.rs:10-96 (prover side).  SRS = g * tau^i, i = 0..max_degree."""

    def __init__(self, max_degree, g1, tau):
        self.max_degree = max_degree
        self.g1 = g1
        self.tau = tau % FR
        self.g1_points = []
        t = 1
        for _ in range(max_degree + 1):  # :44-47
            self.g1_points.append(g1_mul(g1, t))
            t = t * self.tau % FR

    def commit(self, poly):  # :61-73
        assert len(poly) <= self.max_degree + 1, "Polynomial degree exceeds max degree"
        return msm_naive(self.g1_points, poly)

    def open(self, poly, x):  # :75-96
        p = trim(poly)
        y = poly_eval(p, x)
        # (p - y) / (X - x): synthetic division
        num = poly_add(p, trim([(-y) % FR]))
        q = [0] * max(len(num) - 1, 0)
        carry = 0
        for i in reversed(range(1, len(num))):
            carry = (num[i] + carry * x) % FR
            q[i - 1] = carry
        q = trim(q)
        return x % FR, y, self.commit(q)


# ---------------------------------------------------------------- IPA S polynomial / MLPCS open
def compute_s_polynomial(p1, p2):
    """pcs/src/ipa.rs:122-157."""
    L = max(len(p1), len(p2))
    a = [x % FR for x in p1] + [0] * (L - len(p1))
    b = [x % FR for x in p2] + [0] * (L - len(p2))
    h = poly_add(poly_mul(trim(a), trim(b[::-1])), poly_mul(trim(a[::-1]), trim(b)))
    h = h + [0] * (2 * L - 1 - len(h))
    return trim(h[len(h) // 2 + 1:])


def eval_pr(r, x):  # pcs/src/mlpcs.rs:52-63
    res, xp = 1, x % FR
    for ri in r:
        res = res * ((ri * xp + 1 - ri) % FR) % FR
        xp = xp * xp % FR
    return res


def compute_pr(r):
    """pcs/src/mlpcs.rs:68-78.  Coefficient j = prod_i (r_i if bit_i(j) else 1-r_i); the reference gets
    the same numbers via evaluate-on-domain + IFFT.  Trailing zeros trimmed (DensePolynomial)."""
    return trim(fast_eq_eval_hypercube(len(r), list(r)))


def mlpcs_open(kzg: KZG, poly, eval_point, tr: Transcript):
    """pcs/src/mlpcs.rs:83-124."""
    pr = compute_pr(eval_point)
    evaluation = sum(a * b for a, b in zip(poly, pr)) % FR  # :91-94
    s = compute_s_polynomial(list(poly), pr)  # :96
    s_comm = kzg.commit(s)  # :97
    tr.append_fr_vec(eval_point)  # :100
    tr.append_fr(evaluation)  # :101
    tr.append_g1(s_comm)  # :102
    r = tr.draw_field_element()  # :105
    r_inv = fr_inv(r)  # :107
    return dict(
        evaluation_point=[x % FR for x in eval_point], evaluation=evaluation, s_comm=s_comm,
        poly_opening=kzg.open(poly, r), poly_opening_inv=kzg.open(poly, r_inv),
        s_opening=kzg.open(s, r), s_opening_inv=kzg.open(s, r_inv),
    )
