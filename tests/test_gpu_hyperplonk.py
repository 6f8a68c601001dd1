"""GPU parity for the callers of the hot path: logup multiset / permutation check and the
HyperPlonk prove driver, against the Python restatement (oracle/pyref.py) on the reference's integration-test
circuits (hyperplonk/tests/test_basic_proof.rs:17-196) and a larger Fibonacci trace.  Every commitment, round
polynomial, opening and the final Fiat-Shamir state must match."""
import numpy as np
import pytest

import quill_zkvm_b200 as q
from oracle import coracle as co
from oracle import fastkzg
from oracle import pyref as py
from quill_zkvm_b200 import hyperplonk as hp
from tests import util

pytestmark = pytest.mark.gpu
FR = py.FR
GEN = py.g1_mul(py.G1_GEN, 7)
TAU = 0x1234567890ABCDEF1234567890ABCDEF


def to_q(e):
    """pyref tuple expression -> product VirtualPolyExpr"""
    if e[0] == "in":
        return q.VirtualPolyExpr.Input(e[1])
    if e[0] == "const":
        return hp.Const(e[1])
    l, r = to_q(e[1]), to_q(e[2])
    return l + r if e[0] == "add" else l * r


def to_product_circuit(c: py.TransitionCircuit) -> hp.TransitionCircuit:
    out = hp.TransitionCircuit(c.num_rows)
    out.num_columns = c.num_columns
    out.state_cells = list(c.state_cells)
    out.recurring_constraints = [to_q(e) for e in c.recurring]
    out.boundary_constraints = [(row, to_q(e)) for row, e in c.boundary]
    return out


def same_point(a, b):
    return (a is None and not np.asarray(b).any()) or (a is not None and co.g1_from_bytes(b) == a)


def same_opening(got: q.MLEvalProof, want: dict):
    assert co.from_mont(got.evaluation)[0] == want["evaluation"]
    assert same_point(want["s_comm"], got.s_comm)
    for g, key in ((got.poly_opening, "poly_opening"), (got.poly_opening_inv, "poly_opening_inv"),
                   (got.s_opening, "s_opening"), (got.s_opening_inv, "s_opening_inv")):
        x, y, pr = want[key]
        assert co.from_mont(g.x)[0] == x and co.from_mont(g.y)[0] == y and same_point(pr, g.proof)


def run_both(ctx, circuits_py, witnesses):
    max_degree = max(c.num_cols() * c.num_rows for c in circuits_py)
    fastkzg.install_fast_s_polynomial()
    okzg = fastkzg.FastKZG(max_degree, GEN, TAU)
    want = py.hyperplonk_prove(circuits_py, witnesses, okzg)
    kzg = q.KZG.trusted_setup(ctx, max_degree, co.g1_to_bytes(GEN), co.fr1(TAU))
    prover = hp.HyperPlonk.preprocess(ctx, [to_product_circuit(c) for c in circuits_py], kzg)
    got = prover.prove(kzg, [[co.to_mont(col) for col in w] for w in witnesses])
    kzg.srs.free()
    assert got.transcript_state.hex() == want["state_end"]
    for com, wcom in zip(got.witness_commitment, want["witness_commitment"]):
        assert same_point(wcom, com)
    for tp, wp in zip(got.trace_proofs, want["trace_proofs"]):
        zc = tp.zero_check_proof.sumcheck_proof
        assert [co.from_mont(p) for p in zc.r_polys] == wp["zc_polys"]
        ms, wms = tp.permutation_check_proof.multiset_equality_proof, wp["permutation"]
        assert same_point(wms["denom_left_commitment"], ms.denom_left_commitment)
        assert same_point(wms["denom_right_commitment"], ms.denom_right_commitment)
        assert [co.from_mont(p) for p in ms.sumcheck_proof.r_polys] == wms["r_polys"]
        same_opening(ms.opening_proof_denom_left, wms["opening_proof_denom_left"])
        same_opening(ms.opening_proof_denom_right, wms["opening_proof_denom_right"])
        for g, w in zip(tp.openings_zero_check, wp["openings_zero_check"]):
            same_opening(g, w)
        for g, w in zip(tp.openings_public, wp["openings_public"]):
            same_opening(g, w)
        same_opening(tp.opening_id, wp["opening_id"])
        same_opening(tp.opening_permutation, wp["opening_permutation"])
        same_opening(tp.opening_permutation_trace, wp["opening_permutation_trace"])
    return got, want


def test_hyperplonk_fibonacci(ctx):
    """test_hyperplonk_proof (test_basic_proof.rs:137-164): Fibonacci, 8 rows x 4 columns"""
    c, w = py.fibonacci_circuit_and_trace()
    got, want = run_both(ctx, [c], [w])
    # the zero-check claim is h(column evaluations) at the point: constraints hold, so the opened column values satisfy it
    tp = got.trace_proofs[0]
    assert len(tp.openings_zero_check) == 4 and len(tp.openings_public) == 2


def test_hyperplonk_multitrace(ctx):
    """test_hyperplonk_proof_multitrace (test_basic_proof.rs:166-196): Fibonacci + modified Fibonacci (5 columns padded to 8)"""
    c1, w1 = py.fibonacci_circuit_and_trace()
    c2, w2 = py.modified_fibonacci_circuit_and_trace()
    run_both(ctx, [c1, c2], [w1, w2])


def test_hyperplonk_fibonacci_512_rows(ctx):
    c, w = py.fibonacci_circuit_and_trace(512)
    run_both(ctx, [c], [w])


def test_multiset_subset_mode(ctx):
    """multiset_check.rs:67-95 (Subset mode with multiplicities), 5 variables as in the reference's own test"""
    import random
    rnd = random.Random(4)
    n = 5
    table = [rnd.randrange(1, 1000) for _ in range(1 << n)]
    # left = table entries looked up with repetition; multiplicities count how often each table entry is used
    picks = [rnd.randrange(1 << n) for _ in range(1 << n)]
    left = [table[i] for i in picks]
    mult = [picks.count(i) for i in range(1 << n)]
    fastkzg.install_fast_s_polynomial()
    okzg = fastkzg.FastKZG(1 << n, GEN, TAU)
    tabs = [left, table, mult]
    wtr = py.Transcript(b"subset")
    want, wpoint = py.multiset_prove([list(t) for t in tabs], n, py.e_in(0), py.e_in(1), wtr, okzg, multiplicities=py.e_in(2))
    kzg = q.KZG.trusted_setup(ctx, 1 << n, co.g1_to_bytes(GEN), co.fr1(TAU))
    store = q.VirtualPolynomialStore(n)
    for t in tabs:
        store.allocate_polynomial(co.to_mont(t))
    hl, hr, m = (store.new_virtual_from_input(i) for i in range(3))
    tr = q.Transcript(b"subset", ctx)
    got, point = hp.MultisetEqualityProof.prove(ctx, store, hl, hr, tr, kzg, multiplicities=m)
    kzg.srs.free()
    assert co.from_mont(point) == wpoint and tr.state.tobytes().hex() == wtr.state.hex()
    assert [co.from_mont(p) for p in got.sumcheck_proof.r_polys] == want["r_polys"]
    same_opening(got.opening_proof_denom_left, want["opening_proof_denom_left"])
    # the logup identity holds: sum_x 1/(gamma+left) - mult/(gamma+table) = 0, so the verifier accepts claimed sum 0
    mc = q._lib.QZ_MAX_ROUND_COEFFS
    coeffs = np.zeros((n, mc, 32), dtype=np.uint8)
    for j, p in enumerate(got.sumcheck_proof.r_polys):
        coeffs[j, : p.shape[0]] = p
    c0 = co.from_mont(got.sumcheck_proof.r_polys[0])
    assert (2 * (c0[0] if c0 else 0) + sum(c0[1:])) % FR == 0
