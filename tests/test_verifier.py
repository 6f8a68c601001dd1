"""The oracle's restatement of the reference's VERIFIER side (oracle/verifier.py + oracle/pairing.hpp),
exercised the way the reference's own tests are written: prove -> verify -> accept, and a tampered twin -> reject
(SURVEY section 4).  CPU only: prover = oracle/pyref.py; the GPU tests run the same verifier on the product's proofs."""
import copy
import random

import pytest

from oracle import coracle as co
from oracle import fastkzg
from oracle import pyref as py
from oracle import verifier as vf

FR = py.FR
GEN = py.g1_mul(py.G1_GEN, 7)
TAU = 0x1234567890ABCDEF1234567890ABCDEF


@pytest.fixture(scope="module")
def vk():
    fastkzg.install_fast_s_polynomial()
    return vf.VerifierKey(GEN, TAU, g2_scalar=11)


def test_g2_generator_and_pairing_bilinearity():
    g2 = co.g2_generator()
    assert co.g2_on_curve(g2)
    assert co.g2_mul(g2, co.fr1(FR - 1)).any() and not co.g2_add(co.g2_mul(g2, co.fr1(FR - 1)), g2).any()  # order r
    g1 = co.g1_to_bytes(py.G1_GEN)
    one, e = co.pairing_product([g1], [g2])
    assert not one  # non-degenerate
    a, b = 0x1F3A5C7E9B2D4F6081, 0xA1B2C3D4E5F60718293A
    _, e_ab = co.pairing_product([co.g1_mul(g1, co.fr1(a))], [co.g2_mul(g2, co.fr1(b))])
    assert e_ab == co.pairing_product([co.g1_mul(g1, co.fr1(a * b % FR))], [g2])[1]
    assert e_ab == co.pairing_product([g1], [co.g2_mul(g2, co.fr1(a * b % FR))])[1]
    assert e_ab != e
    # e(aP, Q) e(-P, aQ) = 1, and the pairing with the identity is one
    assert co.pairing_product([co.g1_mul(g1, co.fr1(a)), co.g1_to_bytes(py.g1_neg(py.G1_GEN))],
                              [g2, co.g2_mul(g2, co.fr1(a))])[0]
    assert co.pairing_product([co.g1_to_bytes(None)], [g2])[0]


def test_kzg_reference_test(vk):
    """test_kzg (pcs/src/kzg.rs:119-151): SRS degree 4, p = 2 + x + 3x^2 opened at 5; y + 1 rejected"""
    kzg = py.KZG(4, GEN, TAU)
    p = [2, 1, 3]
    c = kzg.commit(p)
    x, y, proof = kzg.open(p, 5)
    assert y == 2 + 5 + 3 * 25
    assert vf.kzg_verify(vk, c, (x, y, proof))
    assert not vf.kzg_verify(vk, c, (x, (y + 1) % FR, proof))
    assert not vf.kzg_verify(vk, c, (x + 1, y, proof))
    assert not vf.kzg_verify(vk, py.g1_add(c, GEN), (x, y, proof))


def _mlpcs_roundtrip(vk, poly, point, srs_degree):
    kzg = fastkzg.FastKZG(srs_degree, GEN, TAU)
    com = kzg.commit(poly)
    proof = py.mlpcs_open(kzg, poly, point, py.Transcript(b"mlpcs_test"))
    assert vf.mlpcs_verify(vk, com, proof, py.Transcript(b"mlpcs_test"))
    bad = dict(proof, evaluation=(proof["evaluation"] + 1) % FR)
    assert not vf.mlpcs_verify(vk, com, bad, py.Transcript(b"mlpcs_test"))
    assert not vf.mlpcs_verify(vk, com, proof, py.Transcript(b"another domain"))
    return proof


def test_mlpcs_reference_tests(vk):
    """test_mlpcs_proof / _zero_opening / _zero_one_opening / _degree_bound (pcs/src/mlpcs.rs:245-474)"""
    rnd = random.Random(5)
    poly = [rnd.randrange(FR) for _ in range(32)]
    point = [rnd.randrange(FR) for _ in range(5)]
    proof = _mlpcs_roundtrip(vk, poly, point, 32)
    assert proof["evaluation"] == py.mle_evaluate(poly, point)  # mlpcs.rs:283-285
    poly8 = [rnd.randrange(FR) for _ in range(8)]
    assert _mlpcs_roundtrip(vk, poly8, [0, 0, 0], 8)["evaluation"] == poly8[0]  # mlpcs.rs:321-356
    assert _mlpcs_roundtrip(vk, poly8, [0, 1, 0], 8)["evaluation"] == poly8[2]  # mlpcs.rs:358-393
    # fewer variables than log2(len): the opening evaluates the 2^3-entry prefix (mlpcs.rs:395-429)
    point3 = [rnd.randrange(FR) for _ in range(3)]
    assert _mlpcs_roundtrip(vk, poly, point3, 128)["evaluation"] == py.mle_evaluate(poly[:8], point3)


def test_zerocheck_verify_and_false_statement():
    """test_zerocheck_proof / _not_zero (zerocheck.rs:85-211): g1 = i, g2 = i^2, h = g1^2 - g2; then g2[3] += 1"""
    n = 3
    g1 = list(range(8))
    g2 = [i * i for i in range(8)]
    h = py.e_sub(py.e_mul(py.e_in(0), py.e_in(0)), py.e_in(1))
    polys, point, ev, _z = py.zerocheck_prove(n, [g1, g2], h, py.Transcript(b"zerocheck_test"))
    vpoint, vev = vf.zerocheck_verify(n, polys, py.Transcript(b"zerocheck_test"))
    assert (vpoint, vev) == (point, ev)
    assert ev == (py.mle_evaluate(g1, point) ** 2 - py.mle_evaluate(g2, point)) % FR  # zerocheck.rs:142-158
    g2[3] += 1
    polys, point, ev, _z = py.zerocheck_prove(n, [g1, g2], h, py.Transcript(b"zerocheck_test"))
    with pytest.raises(ValueError):
        vf.zerocheck_verify(n, polys, py.Transcript(b"zerocheck_test"))


def _multiset_case(vk, left, right, mult=None, n=5):
    kzg = fastkzg.FastKZG(1 << n, GEN, TAU)
    tabs = [left, right] + ([mult] if mult is not None else [])
    m_expr = py.e_in(2) if mult is not None else None
    proof, point = py.multiset_prove([list(t) for t in tabs], n, py.e_in(0), py.e_in(1), py.Transcript(b"multiset"),
                                     kzg, multiplicities=m_expr)
    claims = [(point, py.mle_evaluate(t, point)) for t in tabs]
    vf.multiset_verify(proof, n, py.Transcript(b"multiset"), vk, claims[0], claims[1],
                       multiplicities_eval=claims[2] if mult is not None else None)


def test_multiset_equality_and_subset(vk):
    """multiset_check.rs:310-636: shuffled copy accepted, one changed entry rejected; Subset mode with multiplicities"""
    rnd = random.Random(9)
    n = 5
    left = [rnd.randrange(FR) for _ in range(1 << n)]
    right = list(left)
    rnd.shuffle(right)
    _multiset_case(vk, left, right)
    bad = list(right)
    bad[0] = (bad[0] + 1) % FR
    with pytest.raises(ValueError):
        _multiset_case(vk, left, bad)
    table = [rnd.randrange(1, 1000) for _ in range(1 << n)]
    picks = [rnd.randrange(1 << n) for _ in range(1 << n)]
    looked_up = [table[i] for i in picks]
    mult = [picks.count(i) for i in range(1 << n)]
    _multiset_case(vk, looked_up, table, mult)
    bad_mult = list(mult)
    bad_mult[0] += 1
    with pytest.raises(ValueError):
        _multiset_case(vk, looked_up, table, bad_mult)


def test_permutation_check(vk):
    """permutation_check.rs:106-332: a permuted copy is accepted under its permutation; with two VALUES swapped the
    multisets are still equal but the permutation check rejects (:251-253)"""
    rnd = random.Random(13)
    n = 5
    N = 1 << n
    kzg = fastkzg.FastKZG(N, GEN, TAU)
    left = [rnd.randrange(FR) for _ in range(N)]
    sigma = list(range(N))
    rnd.shuffle(sigma)
    right = [left[sigma[i]] for i in range(N)]
    ids = [i + 1 for i in range(N)]
    perm = [sigma[i] + 1 for i in range(N)]

    def run(right_values):
        proof, point = py.permutation_prove([list(left), list(right_values)], n, py.e_in(0), py.e_in(1), ids, perm,
                                            py.Transcript(b"perm"), kzg)
        ev = lambda t: (point, py.mle_evaluate(t, point))  # noqa: E731
        vf.permutation_verify(proof, n, py.Transcript(b"perm"), vk, ev(left), ev(right_values), ev(ids), ev(perm))

    run(right)
    swapped = list(right)
    swapped[0], swapped[1] = swapped[1], swapped[0]
    with pytest.raises(ValueError):
        run(swapped)


def test_hyperplonk_fibonacci_prove_verify(vk):
    """test_hyperplonk_proof (hyperplonk/tests/test_basic_proof.rs:137-164) + tampered twins"""
    c, w = py.fibonacci_circuit_and_trace()
    kzg = fastkzg.FastKZG(c.num_cols() * c.num_rows, GEN, TAU)
    proof = py.hyperplonk_prove([c], [w], kzg)
    tvk = vf.hyperplonk_vk([c], kzg)
    assert vf.hyperplonk_verify(proof, tvk, vk) == proof["state_end"]
    bad = copy.deepcopy(proof)
    bad["trace_proofs"][0]["zc_polys"][1][0] = (bad["trace_proofs"][0]["zc_polys"][1][0] + 1) % FR
    with pytest.raises(ValueError):
        vf.hyperplonk_verify(bad, tvk, vk)
    bad = copy.deepcopy(proof)
    bad["witness_commitment"][0] = py.g1_add(bad["witness_commitment"][0], GEN)
    with pytest.raises(ValueError):
        vf.hyperplonk_verify(bad, tvk, vk)
    # a witness that breaks the recurrence: the prover's own constraint check (circuit.rs / proof.rs:262) panics
    w_bad = [list(col) for col in w]
    w_bad[2][3] = (w_bad[2][3] + 1) % FR
    with pytest.raises(AssertionError):
        py.hyperplonk_prove([c], [w_bad], kzg)
