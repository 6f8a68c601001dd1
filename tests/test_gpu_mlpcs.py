"""GPU parity: MultilinearPCS::open (MLEvalProof::prove, pcs/src/mlpcs.rs:83-124) and the S polynomial
(pcs/src/ipa.rs:122-157) against the oracle, bit-exact, on the reference's test shapes (mlpcs.rs:245-474,
ipa.rs:214-298) and larger seeded ones."""
import os
import random

import numpy as np
import pytest

import quill_zkvm_b200 as q
from oracle import coracle as co
from oracle import pyref as py
from tests import util

pytestmark = pytest.mark.gpu
FR = py.FR
GEN = py.g1_mul(py.G1_GEN, 7)
TAU = 0x1234567890ABCDEF1234567890ABCDEF
NCPU = os.cpu_count() or 1


@pytest.fixture(scope="module")
def kzg(ctx):
    k = q.KZG.trusted_setup(ctx, (1 << 17) - 1, co.g1_to_bytes(GEN), co.fr1(TAU))
    yield k
    k.srs.free()


def _trim(a):
    n = a.shape[0]
    while n and not a[n - 1].any():
        n -= 1
    return a[:n]


@pytest.mark.parametrize("n1,n2", [(1, 1), (2, 2), (3, 3), (3, 2), (2, 5), (5, 9), (64, 64), (1000, 1000), (4097, 300),
                                   (1 << 15, 1 << 15), (3000, 2049), (1 << 19, 1 << 19), (600000, 1000),
                                   ((1 << 20) + 5, 300000)])
def test_s_polynomial_vs_oracle(ctx, kzg, n1, n2):
    """NTT sizes 2^1 .. 2^22: one pass, two passes and three passes of 10+10, 7+6, 7+7+7 and 8+7+7 fused stages"""
    a, b = util.rand_fr(n1, n1), util.rand_fr(n2, 7 * n2 + 1)
    got = kzg.compute_s_polynomial(a, b)
    want = co.compute_s_polynomial(a, b)
    assert got.shape[0] == max(n1, n2) - 1
    assert np.array_equal(_trim(got), want)


def test_s_polynomial_reference_answers(ctx, kzg):
    """ipa.rs:214-298: [1,2,3].[4,5,6] and the mismatched-degree case"""
    assert co.from_mont(kzg.compute_s_polynomial(co.to_mont([1, 2, 3]), co.to_mont([4, 5, 6]))) == py.compute_s_polynomial([1, 2, 3], [4, 5, 6])
    got = co.from_mont(kzg.compute_s_polynomial(co.to_mont([1, 2, 3]), co.to_mont([4, 5])))
    assert got == py.compute_s_polynomial([1, 2, 3], [4, 5])
    # sparse inputs whose S has trailing zeros
    a = co.to_mont([1, 0, 0, 0, 0, 0, 0, 0])
    b = co.to_mont([0, 0, 0, 0, 0, 1])
    assert np.array_equal(_trim(kzg.compute_s_polynomial(a, b)), co.compute_s_polynomial(a, b))


def _check_open(ctx, kzg, poly, point, domain=b"mlpcs"):
    tr = q.Transcript(domain, ctx)
    pf = kzg.open_multilinear(poly, point, tr)
    st = co.transcript_new(domain)
    n_srs = len(kzg.srs)
    want = co.mlpcs_open(kzg.srs.download(0, min(n_srs, max(poly.shape[0], 1 << point.shape[0]) + 1)), poly, point, st,
                         threads=NCPU)
    assert np.array_equal(pf.evaluation, want["evaluation"])
    assert np.array_equal(pf.s_comm, want["s_comm"])
    for got, (x, y, proof) in zip([pf.poly_opening, pf.poly_opening_inv, pf.s_opening, pf.s_opening_inv], want["openings"]):
        assert np.array_equal(got.x, x) and np.array_equal(got.y, y) and np.array_equal(got.proof, proof)
    assert tr.state.tobytes() == st.tobytes()
    return pf


def test_reference_mlpcs_tests(ctx, kzg):
    rnd = random.Random(1)
    # test_mlpcs_proof (mlpcs.rs:245-319): 5 variables, evaluation == DenseMultilinearExtension::evaluate
    poly = util.rand_fr(32, 5)
    point = util.rand_fr(5, 6)
    pf = _check_open(ctx, kzg, poly, point)
    assert np.array_equal(pf.evaluation, co.mle_evaluate(poly, point))
    # test_mlpcs_zero_opening / _zero_one_opening (mlpcs.rs:321-393): degenerate P_r (a monomial)
    poly = util.rand_fr(8, 7)
    for pt in ([0, 0, 0], [0, 1, 0], [1, 1, 1], [1, 0, 1]):
        pf = _check_open(ctx, kzg, poly, co.to_mont(pt))
        idx = sum(b << i for i, b in enumerate(pt))
        assert np.array_equal(pf.evaluation, poly[idx])
    # test_mlpcs_degree_bound (mlpcs.rs:395-474): a 2^5 polynomial opened at a 3-variable point evaluates the 2^3 prefix
    poly = util.rand_fr(32, 8)
    point = util.rand_fr(3, 9)
    pf = _check_open(ctx, kzg, poly, point)
    assert np.array_equal(pf.evaluation, co.mle_evaluate(poly[:8], point))
    # a point with more variables than the polynomial has entries (P_r longer than the polynomial)
    _check_open(ctx, kzg, util.rand_fr(5, 10), util.rand_fr(4, 11))


@pytest.mark.parametrize("n", [1, 10, 14, 16])
def test_mlpcs_open_sizes(ctx, kzg, n):
    poly = util.rand_fr(1 << n, 100 + n)
    point = util.rand_fr(n, 200 + n)
    pf = _check_open(ctx, kzg, poly, point, b"test_pcs_interface")
    assert np.array_equal(pf.evaluation, co.mle_evaluate(poly, point))


def test_mlpcs_degree_error(ctx):
    small = q.KZG.trusted_setup(ctx, 7, co.g1_to_bytes(GEN), co.fr1(TAU))
    with pytest.raises(AssertionError):
        small.open_multilinear(util.rand_fr(32, 1), util.rand_fr(5, 2), q.Transcript(b"x", ctx))
    small.srs.free()


@pytest.mark.parametrize("n,nv", [(5, 5), (12, 12), (5, 3), (3, 5), (0, 0)])
def test_split_open_equals_fused(ctx, kzg, n, nv):
    """qz_mlpcs_open_begin + the caller's transcript schedule (mlpcs.rs:100-105) + qz_mlpcs_open_finish is the fused
    qz_mlpcs_open: the split that lets independent openings be spread over GPUs"""
    poly = util.rand_fr(1 << n, 300 + n)
    point = util.rand_fr(nv, 400 + nv)
    tr1 = q.Transcript(b"split", ctx)
    fused = kzg.open_multilinear(poly, point, tr1)
    tr2 = q.Transcript(b"split", ctx)
    pend = kzg.open_multilinear_begin(poly, point)
    tr2.append_fr_vec(point)
    tr2.append_fr(pend.evaluation)
    tr2.append_g1(pend.s_comm)
    r = tr2.draw_field_element()
    split = kzg.open_multilinear_finish(pend, r)
    assert tr1.state.tobytes() == tr2.state.tobytes()
    assert np.array_equal(fused.evaluation, split.evaluation) and np.array_equal(fused.s_comm, split.s_comm)
    for a, b in zip((fused.poly_opening, fused.poly_opening_inv, fused.s_opening, fused.s_opening_inv),
                    (split.poly_opening, split.poly_opening_inv, split.s_opening, split.s_opening_inv)):
        assert np.array_equal(a.x, b.x) and np.array_equal(a.y, b.y) and np.array_equal(a.proof, b.proof)
