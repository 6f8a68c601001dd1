"""GPU parity at BASELINE.json's full sizes for the widened rows: config 4 (multilinear PCS commit + open of a
2^22-entry MLE, pcs/src/mlpcs.rs:83-124) and config 5 (HyperPlonk::prove over two 2^20-row transition-circuit traces,
hyperplonk/src/proof/proof.rs:239-301).  The oracle's provers need minutes at these sizes, so every output is pinned
through closed forms on the tau-power SRS and the reference's own acceptance criterion (verify -> accept, tampered ->
reject) instead:

  commitment       C      = p(tau) * g                                   (kzg.rs:61-73 on srs[i] = tau^i g)
  evaluation       v      = MLE(poly)(point)                             (mlpcs.rs:283-285)
  S polynomial     f(u) P_r(1/u) + f(1/u) P_r(u) = 2 v + u S(u) + S(1/u) / u  at a random u   (ipa.rs:122-157)
  s_comm                  = S(tau) * g
  challenge        r      = transcript(point, v, s_comm)                 (mlpcs.rs:100-105)
  each KZG opening (x, y, pi): x in {r, 1/r}, y = p(x), pi = ((p(tau) - y) / (tau - x)) * g      (kzg.rs:75-96)
  verifier         MLEvalProof::verify / HyperPlonkProof::verify accept  (mlpcs.rs:126-161, proof.rs:493-523)
"""
import copy
import os

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


def horner(poly_mont: np.ndarray, x: int) -> int:
    """p(x) by the C++ oracle's KZG::open remainder (kzg.rs:78)"""
    if poly_mont.shape[0] == 0:
        return 0
    y, _ = co.kzg_open_quotient(np.ascontiguousarray(poly_mont), co.fr1(x))
    return co.from_mont(y)[0]


def closed_form_opening(poly_mont, p_tau, opening, want_x):
    x, y, proof = util.kzg_opening_py(opening)
    assert x == want_x
    assert y == horner(poly_mont, x)
    k = (p_tau - y) * pow((TAU - x) % FR, -1, FR) % FR
    assert proof == (py.g1_mul(GEN, k) if k else None)


@pytest.mark.parametrize("n", [18, 22])
def test_mlpcs_commit_open_config4(ctx, n):
    """BASELINE config 4: commit + open of a 2^22-entry MLE (2^18 first: the same checks where the oracle's prover is
    also affordable, compared bit for bit)"""
    N = 1 << n
    kzg = q.KZG.trusted_setup(ctx, N, co.g1_to_bytes(GEN), co.fr1(TAU))
    kzg.precompute()
    poly_dev = ctx.random_fr(N, 4200 + n)
    poly = poly_dev.download().reshape(-1, 32)
    point = util.rand_fr(n, 4300 + n)
    com = kzg.commit(poly_dev)
    tr = q.Transcript(b"mlpcs_config4", ctx)
    pf = kzg.open_multilinear(poly_dev, point, tr)

    p_tau = horner(poly, TAU)
    assert util.g1_py(com) == py.g1_mul(GEN, p_tau)
    v = co.from_mont(co.mle_evaluate(poly, point))[0]
    assert co.from_mont(pf.evaluation)[0] == v
    # the S polynomial, from the same device routine the opening uses, pinned by its defining identity at a random u
    pr = co.compute_pr(point)
    S = kzg.compute_s_polynomial(poly, pr)
    u = 0x5EED5EED5EED5EED5EED5EED5EED5EED5EED
    ui = pow(u, -1, FR)
    pt = co.from_mont(point)
    lhs = (horner(poly, u) * py.eval_pr(pt, ui) + horner(poly, ui) * py.eval_pr(pt, u)) % FR
    rhs = (2 * v + u * horner(S, u) + ui * horner(S, ui)) % FR
    assert lhs == rhs
    s_tau = horner(S, TAU)
    assert util.g1_py(pf.s_comm) == py.g1_mul(GEN, s_tau)
    # transcript schedule (mlpcs.rs:100-105) replayed on the host from the device's outputs
    want_tr = py.Transcript(b"mlpcs_config4")
    want_tr.append_fr_vec(pt)
    want_tr.append_fr(v)
    want_tr.append_g1(util.g1_py(pf.s_comm))
    r = want_tr.draw_field_element()
    assert tr.state.tobytes().hex() == want_tr.state.hex()
    ri = pow(r, -1, FR)
    closed_form_opening(poly, p_tau, pf.poly_opening, r)
    closed_form_opening(poly, p_tau, pf.poly_opening_inv, ri)
    closed_form_opening(S, s_tau, pf.s_opening, r)
    closed_form_opening(S, s_tau, pf.s_opening_inv, ri)
    # the reference's acceptance criterion, pairings included
    vk = vf.VerifierKey(GEN, TAU, g2_scalar=11)
    pd = util.opening_py(pf)
    assert vf.mlpcs_verify(vk, util.g1_py(com), pd, py.Transcript(b"mlpcs_config4"))
    assert not vf.mlpcs_verify(vk, util.g1_py(com), dict(pd, evaluation=(v + 1) % FR), py.Transcript(b"mlpcs_config4"))
    if n <= 18:  # bit for bit against the oracle's MLEvalProof::prove
        st = co.transcript_new(b"mlpcs_config4")
        want = co.mlpcs_open(kzg.srs.download(0, N + 1), poly, point, st, threads=os.cpu_count() or 1)
        assert np.array_equal(pf.evaluation, want["evaluation"]) and np.array_equal(pf.s_comm, want["s_comm"])
        for got, (x, y, proof) in zip([pf.poly_opening, pf.poly_opening_inv, pf.s_opening, pf.s_opening_inv], want["openings"]):
            assert np.array_equal(got.x, x) and np.array_equal(got.y, y) and np.array_equal(got.proof, proof)
        assert tr.state.tobytes() == st.tobytes()
    poly_dev.free()
    kzg.srs.free()


def _prove_and_verify(ctx, log_rows, gen=GEN):
    rows = 1 << log_rows
    c1, w1 = py.fibonacci_circuit_and_trace(rows)
    c2, w2 = py.modified_fibonacci_circuit_and_trace(rows)
    circuits = [c1, c2]
    max_degree = max(c.num_cols() * c.num_rows for c in circuits)
    kzg = q.KZG.trusted_setup(ctx, max_degree, co.g1_to_bytes(gen), co.fr1(TAU)).precompute()
    prover = hp.HyperPlonk.preprocess(ctx, [to_product_circuit(c) for c in circuits], kzg)

    def to_table(col):  # canonical bytes via Python, Montgomery conversion on the device
        raw = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in col), dtype=np.uint8).reshape(-1, 32)
        return ctx.field_op(0, 4, raw)

    witnesses = [[to_table(col) for col in w] for w in (w1, w2)]
    got = prover.prove(kzg, witnesses)
    again = prover.prove(kzg, witnesses)
    assert bytes(got.transcript_state) == bytes(again.transcript_state)  # deterministic: no race in any kernel
    proof = util.hyperplonk_py(got)
    tvk = util.hyperplonk_vk_py(prover.trace_vks, circuits)
    vk = vf.VerifierKey(gen, TAU, g2_scalar=11)
    assert vf.hyperplonk_verify(proof, tvk, vk) == proof["state_end"]
    # witness commitments in closed form: the full witness is the column-major concatenation (proof.rs:270)
    for wit, com in zip(witnesses, got.witness_commitment):
        full = np.concatenate([np.asarray(col).reshape(-1, 32) for col in wit])
        assert util.g1_py(com) == py.g1_mul(gen, horner(full, TAU))
    bad = copy.deepcopy(proof)
    bad["trace_proofs"][1]["opening_id"]["evaluation"] = (bad["trace_proofs"][1]["opening_id"]["evaluation"] + 1) % FR
    with pytest.raises(ValueError):
        vf.hyperplonk_verify(bad, tvk, vk)
    bad = copy.deepcopy(proof)
    bad["trace_proofs"][0]["zc_polys"][0][1] = (bad["trace_proofs"][0]["zc_polys"][0][1] + 1) % FR
    with pytest.raises(ValueError):
        vf.hyperplonk_verify(bad, tvk, vk)
    kzg.srs.free()
    return proof["state_end"]


def test_hyperplonk_config5_small_rows(ctx):
    _prove_and_verify(ctx, 10)


def test_hyperplonk_config5_2_20_rows(ctx):
    """BASELINE config 5: two traces of 2^20 rows (Fibonacci 4 columns, modified Fibonacci 5 columns padded to 8);
    the final transcript state is the one bench.py prints for the same workload at every GPU count"""
    state = _prove_and_verify(ctx, 20, py.G1_GEN)  # bench.py's SRS: the standard generator (1, 2), same tau
    golden = os.path.join(os.path.dirname(__file__), "golden", "hyperplonk_2_20_state.txt")
    assert state == open(golden).read().strip()


def test_bench_digests_pinned_by_the_oracle(ctx):
    """tests/golden/bench_digests_2_24.json holds the commitment and the final transcript states of bench.py's default
    2^24 workload; bench.py compares every run with it at every GPU count (inputs are seeded by global index), so the
    file is what carries parity into the driver's multi-GPU runs.  Here the file itself is held to the oracle: the
    commitment through the closed form p(tau) * g, the sumcheck and zero-check transcripts through the multi-threaded
    C++ restatement on the very same tables."""
    import json

    import bench

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bench_digests_2_24.json")))
    d, n = g["digests"], g["log_n"]
    ncpu = os.cpu_count() or 1
    # commitment of the bench's scalars on srs[i] = tau^i * (1, 2)
    sc = ctx.random_fr(1 << n, bench.shard_seed(0x5155494C4C, 0))
    y, _ = co.kzg_open_quotient(sc.download().reshape(-1, 32), co.fr1(bench.TAU))
    sc.free()
    want = co.g1_mul(co.g1_to_bytes((1, 2)), y)  # the generator bench.py uses
    assert bytes(want).hex() == d["commitment_xy"]
    assert ctx.g1_serialize(want).hex() == d["commitment_serialized"]
    # the three tables, the true sum, both transcripts
    tabs = []
    for t in range(3):
        b = ctx.random_fr(1 << n, bench.shard_seed(1000 * (t + 1), 0))
        tabs.append(b.download().reshape(-1, 32))
        b.free()
    nodes, consts = util.expr_product(3)
    claimed = co.fr1(int(d["sumcheck_claimed_sum"], 16))
    st = co.transcript_new(b"sumcheck_bench")
    o = co.sumcheck_prove(n, tabs, nodes, consts, claimed, st, max_coeffs=8, threads=ncpu)
    c0 = co.from_mont(o["coeffs"][0][: o["lens"][0]])
    assert (2 * c0[0] + sum(c0[1:])) % FR == int(d["sumcheck_claimed_sum"], 16)  # the claim is the true sum
    assert st.tobytes().hex() == d["sumcheck_final_transcript_state"]
    assert bytes(o["evaluation"]).hex() == d["sumcheck_evaluation"]
    st = co.transcript_new(b"zerocheck_bench")
    co.sumcheck_prove(n, tabs, nodes, consts, None, st, max_coeffs=8, zerocheck=True, threads=ncpu)
    assert st.tobytes().hex() == d["zerocheck_final_transcript_state"]
