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


# ---------------------------------------------------------------- logup multiset / permutation check, HyperPlonk driver
def multiset_prove(store_tables, num_vars, h_left, h_right, tr: Transcript, kzg: KZG, multiplicities=None):
    """hyperplonk/src/piops/multiset_check.rs:28-182.  `store_tables` (list of int tables) is extended in place like the
    reference's `&mut store`.  Returns (proof dict, evaluation point)."""
    n = 1 << num_vars
    gamma = tr.draw_field_element()  # :40
    rows = lambda i: [t[i] for t in store_tables]  # noqa: E731
    left = [fr_inv(gamma + expr_eval_point(h_left, rows(i))) for i in range(n)]  # :43-53
    right = [fr_inv(gamma + expr_eval_point(h_right, rows(i))) for i in range(n)]  # :55-65
    if multiplicities is not None:  # :67-95
        right = [r * expr_eval_point(multiplicities, rows(i)) % FR for i, r in enumerate(right)]
    c_left, c_right = kzg.commit(left), kzg.commit(right)  # :98-99
    tr.append_g1(c_left)
    tr.append_g1(c_right)
    lam = tr.draw_field_element()  # :104-105
    alpha = tr.draw_field_element()
    dl, dr = len(store_tables), len(store_tables) + 1  # :108-109
    store_tables.append(left)
    store_tables.append(right)
    m = multiplicities if multiplicities is not None else e_const(1)
    zc = e_add(e_sub(e_mul(e_in(dl), e_add(e_const(gamma), h_left)), e_const(1)),
               e_mul(e_const(lam), e_sub(e_mul(e_in(dr), e_add(e_const(gamma), h_right)), m)))  # :132-141
    z = [tr.draw_field_element() for _ in range(num_vars)]  # :144-146
    eq = len(store_tables)
    store_tables.append(fast_eq_eval_hypercube(num_vars, z))  # :149-150
    h_hat = e_mul(zc, e_in(eq))  # :152-153
    h_hat = e_mul(h_hat, e_const(alpha))  # :156
    h_hat = e_add(h_hat, e_in(dl))  # :157
    h_hat = e_add(h_hat, e_mul(e_const(FR - 1), e_in(dr)))  # :158
    r_polys, point, _ev = sumcheck_prove(num_vars, store_tables, h_hat, 0, tr)  # :162-163
    o_left = mlpcs_open(kzg, left, point, tr)  # :167-170
    o_right = mlpcs_open(kzg, right, point, tr)
    return dict(denom_left_commitment=c_left, denom_right_commitment=c_right, r_polys=r_polys,
                opening_proof_denom_left=o_left, opening_proof_denom_right=o_right), point


def permutation_prove(store_tables, num_vars, h_left, h_right, id_indices, permutation_indices, tr: Transcript, kzg: KZG):
    """hyperplonk/src/piops/permutation_check.rs:13-58"""
    assert len(id_indices) == 1 << num_vars and len(permutation_indices) == 1 << num_vars
    id_ref, perm_ref = len(store_tables), len(store_tables) + 1
    store_tables.append(list(id_indices))
    store_tables.append(list(permutation_indices))
    alpha = tr.draw_field_element()
    hl = e_add(e_mul(h_left, e_const(alpha)), e_in(id_ref))
    hr = e_add(e_mul(h_right, e_const(alpha)), e_in(perm_ref))
    return multiset_prove(store_tables, num_vars, hl, hr, tr, kzg)


class TransitionCircuit:
    """hyperplonk/src/frontend/transition_circuit.rs:26-151 on Python ints / tuple expressions."""

    def __init__(self, num_rows):
        self.num_columns, self.num_rows = 0, num_rows
        self.state_cells, self.recurring, self.boundary = [], [], []

    def allocate_witness_cell(self):
        self.num_columns += 1
        return self.num_columns - 1

    def allocate_state_cell(self):
        c, n = self.allocate_witness_cell(), self.allocate_witness_cell()
        self.state_cells.append((c, n))
        return c, n

    def num_cols(self):
        p = 1
        while p < self.num_columns:
            p *= 2
        return p

    def public_values(self):
        pub = [[0] * self.num_rows for _ in self.boundary]
        for i, (row, _) in enumerate(self.boundary):
            pub[i][row] = 1
        return pub

    def zero_check_expressions(self):
        return list(self.recurring) + [e_mul(e_in(i + self.num_cols()), c) for i, (_r, c) in enumerate(self.boundary)]

    def permutation(self):
        n = self.num_rows * self.num_cols()
        perm = list(range(n))
        for cur, nxt in self.state_cells:
            for row in range(self.num_rows - 1):
                frm, to = nxt * self.num_rows + row, cur * self.num_rows + row + 1
                perm[frm], perm[to] = to, frm
        return [i + 1 for i in range(n)], [p + 1 for p in perm]

    def check_constraints(self, witness):
        for row in range(self.num_rows):
            vals = [col[row] for col in witness]
            for c in self.recurring:
                assert expr_eval_point(c, vals) == 0, f"recurring constraint violated at row {row}"
        for row, c in self.boundary:
            assert expr_eval_point(c, [col[row] for col in witness]) == 0, f"boundary constraint violated at row {row}"
        for cur, nxt in self.state_cells:
            for row in range(self.num_rows - 1):
                assert witness[nxt][row] == witness[cur][row + 1], "copy constraint violated"


def hyperplonk_prove(circuits, witness_traces, kzg: KZG):
    """hyperplonk/src/proof/proof.rs:63-301 (preprocess_trace's prover key + prove).  Returns a dict with the
    commitments, per-trace proofs and the final transcript state."""
    tr = Transcript(b"hyperplonk_proof")  # :245
    comms, fulls = [], []
    for c, witness in zip(circuits, witness_traces):  # :250-284
        assert len(witness) == c.num_cols()
        c.check_constraints(witness)
        full = [v for col in witness for v in col]
        com = kzg.commit(full)
        tr.append_g1(com)
        comms.append(com)
        fulls.append(full)
    proofs = []
    for c, witness, full in zip(circuits, witness_traces, fulls):  # prove_trace :145-237
        log2_rows, log2_cols = c.num_rows.bit_length() - 1, c.num_cols().bit_length() - 1
        tables = [list(col) for col in witness] + c.public_values()
        alpha = tr.draw_field_element()
        zc_expr = e_const(0)
        for i, e in enumerate(c.zero_check_expressions()):
            zc_expr = e_add(zc_expr, e_mul(e_const(pow(alpha, i, FR)), e))
        zc_polys, zc_point, zc_eval, zc_z = zerocheck_prove(log2_rows, tables, zc_expr, tr)
        ids, perm = c.permutation()
        store2 = [list(full)]
        perm_proof, perm_point = permutation_prove(store2, log2_rows + log2_cols, e_in(0), e_in(0), ids, perm, tr, kzg)
        openings_zc = []
        for col in range(c.num_cols()):
            point = list(zc_point) + [(col >> i) & 1 for i in range(log2_cols)]
            openings_zc.append(mlpcs_open(kzg, full, point, tr))
        openings_pub = [mlpcs_open(kzg, p, zc_point, tr) for p in c.public_values()]
        o_id = mlpcs_open(kzg, ids, perm_point, tr)
        o_perm = mlpcs_open(kzg, perm, perm_point, tr)
        o_trace = mlpcs_open(kzg, full, perm_point, tr)
        proofs.append(dict(zc_polys=zc_polys, zc_point=zc_point, zc_eval=zc_eval, permutation=perm_proof,
                           perm_point=perm_point, openings_zero_check=openings_zc, openings_public=openings_pub,
                           opening_id=o_id, opening_permutation=o_perm, opening_permutation_trace=o_trace))
    return dict(witness_commitment=comms, trace_proofs=proofs, state_end=tr.state.hex())


def fibonacci_circuit_and_trace(num_rows=8):
    """hyperplonk/tests/test_basic_proof.rs:17-52"""
    c = TransitionCircuit(num_rows)
    s1, s2 = c.allocate_state_cell(), c.allocate_state_cell()
    c.boundary.append((0, e_in(s1[0])))
    c.boundary.append((0, e_sub(e_in(s2[0]), e_const(1))))
    c.recurring.append(e_sub(e_in(s2[1]), e_add(e_in(s1[0]), e_in(s2[0]))))
    c.recurring.append(e_sub(e_in(s1[1]), e_in(s2[0])))
    w = [[0] * num_rows for _ in range(c.num_cols())]
    for row in range(num_rows):
        if row == 0:
            w[s1[0]][0], w[s2[0]][0], w[s1[1]][0], w[s2[1]][0] = 0, 1, 1, 1
        else:
            w[s1[0]][row] = w[s1[1]][row - 1]
            w[s2[0]][row] = w[s2[1]][row - 1]
            w[s1[1]][row] = w[s2[0]][row]
            w[s2[1]][row] = (w[s2[0]][row] + w[s1[0]][row]) % FR
    return c, w


def modified_fibonacci_circuit_and_trace(num_rows=8):
    """hyperplonk/tests/test_basic_proof.rs:54-105: f(n) = f(n-1) + f(n-1) * f(n-2), 5 columns padded to 8"""
    c = TransitionCircuit(num_rows)
    s1, s2 = c.allocate_state_cell(), c.allocate_state_cell()
    tmp = c.allocate_witness_cell()
    c.boundary.append((0, e_sub(e_in(s1[0]), e_const(1))))
    c.boundary.append((0, e_sub(e_in(s2[0]), e_const(1))))
    c.recurring.append(e_sub(e_in(tmp), e_mul(e_in(s1[0]), e_in(s2[0]))))
    c.recurring.append(e_sub(e_in(s2[1]), e_add(e_in(s1[0]), e_in(tmp))))
    c.recurring.append(e_sub(e_in(s1[1]), e_in(s2[0])))
    w = [[0] * num_rows for _ in range(c.num_cols())]
    for row in range(num_rows):
        if row == 0:
            w[s1[0]][0], w[s2[0]][0] = 1, 1
        else:
            w[s1[0]][row] = w[s1[1]][row - 1]
            w[s2[0]][row] = w[s2[1]][row - 1]
        w[s1[1]][row] = w[s2[0]][row]
        w[tmp][row] = w[s1[0]][row] * w[s2[0]][row] % FR
        w[s2[1]][row] = (w[s1[0]][row] + w[tmp][row]) % FR
    return c, w
