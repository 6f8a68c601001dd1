"""The product's proofs under the reference's acceptance criterion (every reference test is
prove -> verify -> accept plus a tampered twin -> reject).  Prover = the CUDA path through the C ABI; verifier = the
oracle's restatement of the reference's verifier incl. the BN254 pairing (oracle/verifier.py, oracle/pairing.hpp)."""
import copy
import random

import numpy as np
import pytest

import quill_zkvm_b200 as q
from oracle import coracle as co
from oracle import pyref as py
from oracle import verifier as vf
from quill_zkvm_b200 import hyperplonk as hp
from tests import util
from tests.test_gpu_hyperplonk import to_product_circuit

pytestmark = pytest.mark.gpu
FR = py.FR
GEN = py.g1_mul(py.G1_GEN, 7)
TAU = 0x1234567890ABCDEF1234567890ABCDEF


@pytest.fixture(scope="module")
def vk():
    return vf.VerifierKey(GEN, TAU, g2_scalar=11)


@pytest.fixture(scope="module")
def kzg(ctx):
    k = q.KZG.trusted_setup(ctx, 1 << 13, co.g1_to_bytes(GEN), co.fr1(TAU))
    yield k
    k.srs.free()


def test_kzg_open_verifies(ctx, kzg, vk):
    """test_kzg (pcs/src/kzg.rs:119-151) and a 2^13-coefficient polynomial"""
    for poly, x in ((co.to_mont([2, 1, 3]), 5), (util.rand_fr(1 << 13, 3), 0xDEADBEEF)):
        c = util.g1_py(kzg.commit(poly))
        o = util.kzg_opening_py(kzg.open(poly, co.fr1(x)))
        assert vf.kzg_verify(vk, c, o)
        assert not vf.kzg_verify(vk, c, (o[0], (o[1] + 1) % FR, o[2]))


@pytest.mark.parametrize("n,nv", [(5, 5), (3, 3), (5, 3), (12, 12)])
def test_mlpcs_open_verifies(ctx, kzg, vk, n, nv):
    """test_pcs_interface (hyperplonk/tests/test_basic_proof.rs:107-135) / mlpcs.rs:245-474 shapes"""
    poly = util.rand_fr(1 << n, 40 + n)
    point = util.rand_fr(nv, 50 + nv)
    com = util.g1_py(kzg.commit(poly))
    pf = util.opening_py(kzg.open_multilinear(poly, point, q.Transcript(b"test_pcs_interface", ctx)))
    assert vf.mlpcs_verify(vk, com, pf, py.Transcript(b"test_pcs_interface"))
    assert not vf.mlpcs_verify(vk, com, dict(pf, evaluation=(pf["evaluation"] + 1) % FR), py.Transcript(b"test_pcs_interface"))
    assert not vf.mlpcs_verify(vk, py.g1_add(com, GEN), pf, py.Transcript(b"test_pcs_interface"))


def test_zerocheck_verifies_and_false_statement_rejected(ctx):
    """zerocheck.rs:85-211 at 2^10: g2 = g1^2 pointwise accepted; one changed entry -> the verifier rejects"""
    n = 10
    rnd = random.Random(2)
    g1 = [rnd.randrange(FR) for _ in range(1 << n)]
    g2 = [v * v % FR for v in g1]
    for tamper in (False, True):
        if tamper:
            g2[77] = (g2[77] + 1) % FR
        store = q.VirtualPolynomialStore(n)
        a, b = store.allocate_polynomial(co.to_mont(g1)), store.allocate_polynomial(co.to_mont(g2))
        h = store.new_virtual_from_expr(hp.Sub(q.VirtualPolyExpr.Input(a) * q.VirtualPolyExpr.Input(a), q.VirtualPolyExpr.Input(b)))
        zc, claim = q.ZeroCheckProof.prove(ctx, store, h, q.Transcript(b"zerocheck_test", ctx))
        polys = [co.from_mont(p) if len(p) else [] for p in zc.sumcheck_proof.r_polys]
        if tamper:
            with pytest.raises(ValueError):
                vf.zerocheck_verify(n, polys, py.Transcript(b"zerocheck_test"))
        else:
            point, ev = vf.zerocheck_verify(n, polys, py.Transcript(b"zerocheck_test"))
            assert point == co.from_mont(claim.point) and ev == co.from_mont(claim.evaluation)[0]
            assert ev == (py.mle_evaluate(g1, point) ** 2 - py.mle_evaluate(g2, point)) % FR


def test_multiset_and_permutation_verify(ctx, kzg, vk):
    """multiset_check.rs:310-636 / permutation_check.rs:106-332 at 7 variables (the reference's size)"""
    n = 7
    N = 1 << n
    rnd = random.Random(21)
    left = [rnd.randrange(FR) for _ in range(N)]
    sigma = list(range(N))
    rnd.shuffle(sigma)
    right = [left[s] for s in sigma]
    ids, perm = [i + 1 for i in range(N)], [s + 1 for s in sigma]

    def multiset(right_values):
        store = q.VirtualPolynomialStore(n)
        store.allocate_polynomial(co.to_mont(left))
        store.allocate_polynomial(co.to_mont(right_values))
        hl, hr = store.new_virtual_from_input(0), store.new_virtual_from_input(1)
        proof, point = hp.MultisetEqualityProof.prove(ctx, store, hl, hr, q.Transcript(b"multiset", ctx), kzg)
        pt = co.from_mont(point)
        vf.multiset_verify(util.multiset_py(proof), n, py.Transcript(b"multiset"), vk,
                           (pt, py.mle_evaluate(left, pt)), (pt, py.mle_evaluate(right_values, pt)))

    multiset(right)
    bad = list(right)
    bad[0] = (bad[0] + 1) % FR
    with pytest.raises(ValueError):
        multiset(bad)

    def permutation(right_values):
        store = q.VirtualPolynomialStore(n)
        store.allocate_polynomial(co.to_mont(left))
        store.allocate_polynomial(co.to_mont(right_values))
        hl, hr = store.new_virtual_from_input(0), store.new_virtual_from_input(1)
        proof, point = hp.PermutationCheckProof.prove(ctx, store, hl, hr, co.to_mont(ids), co.to_mont(perm),
                                                      q.Transcript(b"perm", ctx), kzg)
        pt = co.from_mont(point)
        ev = lambda t: (pt, py.mle_evaluate(t, pt))  # noqa: E731
        vf.permutation_verify(util.multiset_py(proof.multiset_equality_proof), n, py.Transcript(b"perm"), vk,
                              ev(left), ev(right_values), ev(ids), ev(perm))

    permutation(right)
    swapped = list(right)
    swapped[0], swapped[1] = swapped[1], swapped[0]
    with pytest.raises(ValueError):
        permutation(swapped)  # multisets equal, permutation wrong (permutation_check.rs:251-253)


def test_hyperplonk_multitrace_verifies(ctx, vk):
    """test_hyperplonk_proof_multitrace (hyperplonk/tests/test_basic_proof.rs:166-196): prove on the device, verify
    with the reference's verifier restated; tampered proofs rejected"""
    c1, w1 = py.fibonacci_circuit_and_trace()
    c2, w2 = py.modified_fibonacci_circuit_and_trace()
    circuits = [c1, c2]
    max_degree = max(c.num_cols() * c.num_rows for c in circuits)
    kz = q.KZG.trusted_setup(ctx, max_degree, co.g1_to_bytes(GEN), co.fr1(TAU))
    prover = hp.HyperPlonk.preprocess(ctx, [to_product_circuit(c) for c in circuits], kz)
    proof = util.hyperplonk_py(prover.prove(kz, [[co.to_mont(col) for col in w] for w in (w1, w2)]))
    kz.srs.free()
    tvk = util.hyperplonk_vk_py(prover.trace_vks, circuits)
    assert vf.hyperplonk_verify(proof, tvk, vk) == proof["state_end"]
    bad = copy.deepcopy(proof)
    bad["trace_proofs"][1]["opening_id"]["evaluation"] = (bad["trace_proofs"][1]["opening_id"]["evaluation"] + 1) % FR
    with pytest.raises(ValueError):
        vf.hyperplonk_verify(bad, tvk, vk)
    bad = copy.deepcopy(proof)
    bad["trace_proofs"][0]["zc_polys"][0][1] = (bad["trace_proofs"][0]["zc_polys"][0][1] + 1) % FR
    with pytest.raises(ValueError):
        vf.hyperplonk_verify(bad, tvk, vk)
