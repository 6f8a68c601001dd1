"""CPU tests: pin the oracle against the committed golden vectors, the `blake3` wheel (the reference's own hash
crate), the reference's known answers, and cross-check the C++ restatement against the Python one."""
import json
import os
import random

import blake3
import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyref as py
from tests import util

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
FR, FQ = py.FR, py.FQ
hx = lambda s: int(s, 16)  # noqa: E731


def test_constants_match_survey():
    assert FR == 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    assert FQ == 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    assert (1 << 256) % FR == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert co.from_mont(co.to_mont([5]))[0] == 5
    # Montgomery form of 1 as the C++ oracle computes it == R mod r (the survey's constant)
    one = co.field_op(0, 4, np.frombuffer((1).to_bytes(32, "little"), dtype=np.uint8).copy())
    assert int.from_bytes(one.tobytes(), "little") == (1 << 256) % FR


@pytest.mark.parametrize("field,mod", [(0, FR), (1, FQ)])
def test_inverse_binary_euclid_matches_fermat_and_python(field, mod):
    """Field::inverse of the oracle is the binary extended Euclid of ark-ff (oracle/field.hpp); it must agree with the
    Fermat power it replaced and with Python's modular inverse, on random values and on the edge values."""
    rnd = random.Random(1234 + field)
    vals = [0, 1, 2, mod - 1, mod - 2, (mod + 1) // 2, 1 << 255 % mod] + [rnd.randrange(mod) for _ in range(200)]
    R = 1 << 256
    mont = np.frombuffer(b"".join((v * R % mod).to_bytes(32, "little") for v in vals), dtype=np.uint8).reshape(-1, 32)
    inv = co.field_op(field, 3, mont)
    fermat = co.field_op(field, 6, mont)
    assert np.array_equal(inv, fermat)
    for v, row in zip(vals, inv):
        got = int.from_bytes(row.tobytes(), "little") * pow(R, -1, mod) % mod
        assert got == (pow(v, -1, mod) if v else 0)


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 1023, 1024, 1025, 2048, 3073, 9000])
def test_blake3_matches_wheel(n):
    rnd = random.Random(n)
    d = bytes(rnd.randrange(256) for _ in range(n))
    for out_len in (32, 48, 200):
        assert co.blake3(d, out_len) == blake3.blake3(d).digest(length=out_len)


def test_blake3_published_vector():
    # hash of the empty input, from the BLAKE3 specification
    assert co.blake3(b"").hex() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"
    assert blake3.blake3(b"").hexdigest() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"


def test_transcript_golden_both_oracles():
    for dom, g in G["transcript"].items():
        d = dom.encode()
        t = py.Transcript(d)
        st = co.transcript_new(d)
        assert t.state.hex() == g["state0"] == st.tobytes().hex()
        f1 = t.draw_field_element()
        c1 = co.from_mont(co.transcript_draw_fr(st))[0]
        assert f1 == c1 == hx(g["fe1"])
        t.append_usize(3), t.append_fr(48), t.append_fr_vec([0, 38, 10])
        co.transcript_append(st, py.ser_usize(3))
        co.transcript_append(st, py.ser_fr(48))
        co.transcript_append(st, py.ser_fr_vec([0, 38, 10]))
        assert t.draw_field_element() == co.from_mont(co.transcript_draw_fr(st))[0] == hx(g["fe2"])


def test_survey_kats():
    """values derived independently by the survey's own restatement."""
    t = py.Transcript(b"sumcheck_test")
    assert t.state.hex() == "b7b7e72aa87d8961ed93921435c4eba49fe9ba757f471fb5f10c294b8a70fdc3"
    assert t.draw_field_element() == 15494200051891961783909833458049727794161400064843832742164888701625601212558
    g = G["sumcheck_test"]
    assert hx(g["claimed_sum"]) == 48
    assert [hx(c) for c in g["r_polys"][0]] == [0, 38, 10]
    assert hx(g["point"][0]) == 0x0cce3203f4be7dd22e70fc8de4b0f9440ba7adf39796527ddab1f5544c93952b
    assert hx(g["evaluation"]) == 0x13a45db4205cbe522d2f3a08060b165a99a405183d5ca9709d7df5314bd75c61
    assert g["state_end"] == "be8b0b460434fe5f8f48f50b38e6eafcfdfc6d69c3d049a9cc75f99196d0939c"


def _cpp_sumcheck(n, tables_int, expr_py, claimed, domain, zerocheck=False):
    nodes, consts = util.expr_from_py(expr_py)
    st = co.transcript_new(domain)
    o = co.sumcheck_prove(n, [co.to_mont(t) for t in tables_int], nodes, consts,
                          co.fr1(claimed) if claimed is not None else None, st, max_coeffs=8, zerocheck=zerocheck)
    polys = [co.from_mont(o["coeffs"][j][: o["lens"][j]]) for j in range(n)]
    return polys, co.from_mont(o["point"]), co.from_mont(o["evaluation"])[0], st.tobytes().hex(), o


def test_sumcheck_reference_test_shape():
    """hyperplonk/src/piops/sumcheck.rs:159-230, both restatements vs the golden file + the reference's own asserts."""
    g = G["sumcheck_test"]
    g1 = [((i >> 0) & 1) + 2 * ((i >> 1) & 1) + 3 * ((i >> 2) & 1) for i in range(8)]
    g2 = [((i >> 0) & 1) * 2 * ((i >> 1) & 1) + 3 * ((i >> 0) & 1) * ((i >> 2) & 1) for i in range(8)]
    h = py.e_mul(py.e_in(0), py.e_in(1))
    polys, point, ev, state, o = _cpp_sumcheck(3, [g1, g2], h, 48, b"sumcheck_test")
    assert polys == [[hx(c) for c in p] for p in g["r_polys"]]
    assert point == [hx(x) for x in g["point"]] and ev == hx(g["evaluation"]) and state == g["state_end"]
    # verifier accepts and agrees (sumcheck.rs:203-214)
    st = co.transcript_new(b"sumcheck_test")
    ok, vpoint, vev = co.sumcheck_verify(3, co.fr1(48), o["coeffs"], o["lens"], st)
    assert ok and co.from_mont(vpoint) == point and co.from_mont(vev)[0] == ev
    # closed form (sumcheck.rs:216-229)
    g1r = (point[0] + 2 * point[1] + 3 * point[2]) % FR
    g2r = (point[0] * 2 * point[1] + 3 * point[0] * point[2]) % FR
    assert g1r * g2r % FR == ev
    # a tampered proof is rejected
    bad = o["coeffs"].copy()
    bad[1, 0] = co.fr1(5)
    ok, _, _ = co.sumcheck_verify(3, co.fr1(48), bad, o["lens"], co.transcript_new(b"sumcheck_test"))
    assert not ok


@pytest.mark.parametrize("name", ["zerocheck_test", "zerocheck_test_not_zero"])
def test_zerocheck_reference_test_shapes(name):
    """hyperplonk/src/piops/zerocheck.rs:85-211 (the prover also runs on a false statement)."""
    g = G[name]
    g1v = list(range(8))
    hz = py.e_sub(py.e_mul(py.e_in(0), py.e_in(0)), py.e_in(1))
    polys, point, ev, state, o = _cpp_sumcheck(3, [g1v, g["g2"]], hz, None, b"zerocheck_test", zerocheck=True)
    assert polys == [[hx(c) for c in p] for p in g["r_polys"]]
    assert point == [hx(x) for x in g["point"]] and ev == hx(g["evaluation"]) and state == g["state_end"]
    assert co.from_mont(o["z"]) == [hx(x) for x in g["z"]]
    if name == "zerocheck_test":  # zerocheck.rs:142-158: claim == h(MLE evals at the point)
        a = py.mle_evaluate(g1v, point)
        b = py.mle_evaluate(g["g2"], point)
        assert (a * a - b) % FR == ev
        assert co.from_mont(co.mle_evaluate(co.to_mont(g1v), co.to_mont(point)))[0] == a
    # the sumcheck verifier accepts only the true statement (claimed sum 0)
    st = co.transcript_new(b"zerocheck_test")
    for _ in range(3):
        co.transcript_draw_fr(st)
    ok, _, _ = co.sumcheck_verify(3, co.fr1(0), o["coeffs"], o["lens"], st)
    assert ok == (name == "zerocheck_test")


def test_seeded_product_and_mixed_golden():
    g = G["product3_n6"]
    tabs = [[hx(x) for x in tb] for tb in g["tables"]]
    h3 = py.e_mul(py.e_mul(py.e_in(0), py.e_in(1)), py.e_in(2))
    polys, point, ev, state, _ = _cpp_sumcheck(6, tabs, h3, hx(g["claimed_sum"]), b"sumcheck_bench")
    assert polys == [[hx(c) for c in p] for p in g["r_polys"]] and state == g["state_end"] and ev == hx(g["evaluation"])
    gm = G["mixed_n6"]
    hm = py.e_add(py.e_sub(py.e_mul(py.e_in(0), py.e_in(1)), py.e_in(3)),
                  py.e_mul(py.e_const(7), py.e_mul(py.e_in(2), py.e_in(2))))
    polys, point, ev, state, _ = _cpp_sumcheck(6, tabs, hm, 123, b"mixed")
    assert polys == [[hx(c) for c in p] for p in gm["r_polys"]] and state == gm["state_end"] and ev == hx(gm["evaluation"])


def test_cpp_vs_python_random_expressions():
    rnd = random.Random(7)
    for trial in range(6):
        n = rnd.randrange(1, 6)
        k = rnd.randrange(1, 4)
        tabs = [[rnd.randrange(FR) for _ in range(1 << n)] for _ in range(k)]

        def rand_expr(depth):
            if depth == 0 or rnd.random() < 0.3:
                return py.e_in(rnd.randrange(k)) if rnd.random() < 0.8 else py.e_const(rnd.randrange(FR))
            f = rnd.choice([py.e_add, py.e_mul, py.e_sub])
            return f(rand_expr(depth - 1), rand_expr(depth - 1))

        h = rand_expr(3)
        t = py.Transcript(b"rand%d" % trial)
        rp, pt, ev = py.sumcheck_prove(n, tabs, h, 99, t)
        polys, point, cev, state, _ = _cpp_sumcheck(n, tabs, h, 99, b"rand%d" % trial)
        assert polys == rp and point == pt and cev == ev and state == t.state.hex()


def test_eq_table_golden_and_definition():
    g = G["eq_n5"]
    pt = [hx(x) for x in g["point"]]
    tab = co.from_mont(co.eq_table(co.to_mont(pt)))
    assert tab == [hx(x) for x in g["table"]]
    for i in range(32):  # eq_eval.rs:61-74
        e = 1
        for j in range(5):
            e = e * (pt[j] if (i >> j) & 1 else (1 - pt[j])) % FR
        assert tab[i] == e
    x = [3, 5, 7, 11, 13]
    assert co.from_mont(co.eq_eval(co.to_mont(x), co.to_mont(pt)))[0] == py.eq_eval(x, pt)


def test_g1_known_points_and_serialization():
    two_g = (1368015179489954701390400359078579693043519447331113978918064868415326638035,
             9918110051302171585080402603319702774565515993150576347155970296011118125764)  # EIP-196 test value of 2*G
    assert py.g1_mul(py.G1_GEN, 2) == two_g
    Gb = co.g1_to_bytes(py.G1_GEN)
    assert co.g1_from_bytes(co.g1_add(Gb, Gb)) == two_g
    assert co.g1_from_bytes(co.g1_mul(Gb, co.fr1(FR - 1))) == py.g1_neg(py.G1_GEN)
    assert co.g1_from_bytes(co.g1_add(Gb, co.g1_to_bytes(py.g1_neg(py.G1_GEN)))) is None
    for k in (1, 2, 3, 12345, FR - 2):
        p = py.g1_mul(py.G1_GEN, k)
        assert co.g1_serialize(co.g1_to_bytes(p)) == py.ser_g1(p)
    assert co.g1_serialize(co.g1_to_bytes(None)) == py.ser_g1(None) == bytes(63) + b"\x40"


def test_kzg_golden_commit_open():
    g = G["kzg_test"]
    gen = (hx(g["g"][0]), hx(g["g"][1]))
    srs = co.srs_generate(co.g1_to_bytes(gen), co.fr1(hx(g["tau"])), 5, threads=2)
    assert [co.g1_from_bytes(srs[i]) for i in range(5)] == [(hx(p[0]), hx(p[1])) for p in g["srs"]]
    poly = co.to_mont([2, 1, 3])
    com = co.msm(srs, poly, mode=1)
    assert co.g1_from_bytes(com) == (hx(g["commitment"][0]), hx(g["commitment"][1]))
    assert co.g1_serialize(com).hex() == g["commitment_bytes"]
    assert co.g1_from_bytes(co.msm(srs, poly, mode=0)) == co.g1_from_bytes(com)
    y, q = co.kzg_open_quotient(poly, co.fr1(5))
    assert co.from_mont(y)[0] == hx(g["y"]) == 82
    assert co.g1_from_bytes(co.msm(srs, q)) == (hx(g["proof"][0]), hx(g["proof"][1]))
    # KZG::commit in the reference's own shape (per-call normalisation) gives the same point
    pt, _, _ = co.kzg_commit_reference_shape(srs, poly)
    assert co.g1_from_bytes(pt) == co.g1_from_bytes(com)
    with pytest.raises(AssertionError):
        co.kzg_commit_reference_shape(srs[:2], poly)  # "Polynomial degree exceeds max degree" (kzg.rs:62-65)


def test_msm_golden_and_edge_cases():
    g = G["msm64"]
    k = G["kzg_test"]
    gen = (hx(k["g"][0]), hx(k["g"][1]))
    srs = co.srs_generate(co.g1_to_bytes(gen), co.fr1(hx(k["tau"])), 64, threads=2)
    sc = co.to_mont([hx(s) for s in g["scalars"]])
    want = (hx(g["result"][0]), hx(g["result"][1]))
    assert co.g1_from_bytes(co.msm(srs, sc, mode=1)) == want
    assert co.g1_from_bytes(co.msm(srs, sc, mode=1, threads=3)) == want
    assert co.g1_from_bytes(co.msm(srs, sc, mode=0)) == want
    # msm_unchecked truncates to the shorter input; empty -> identity; P + (-P) -> identity
    assert co.g1_from_bytes(co.msm(srs[:10], sc)) == co.g1_from_bytes(co.msm(srs, sc[:10]))
    assert co.g1_from_bytes(co.msm(srs, sc[:0])) is None
    two = np.stack([srs[3], srs[3]])
    assert co.g1_from_bytes(co.msm(two, co.to_mont([5, FR - 5]))) is None
    # repeated points (doubling inside a bucket)
    rep = np.stack([srs[1]] * 40)
    assert co.g1_from_bytes(co.msm(rep, co.to_mont([3] * 40))) == py.g1_mul(co.g1_from_bytes(srs[1]), 120)


def test_pippenger_vs_naive_random():
    rnd = random.Random(11)
    gen = py.g1_mul(py.G1_GEN, 99)
    srs = co.srs_generate(co.g1_to_bytes(gen), co.fr1(31337), 300, threads=2)
    sc = co.to_mont([rnd.randrange(FR) for _ in range(300)])
    assert co.g1_from_bytes(co.msm(srs, sc, mode=1)) == co.g1_from_bytes(co.msm(srs, sc, mode=0))


def test_pr_and_s_polynomial_reference_answers():
    """pcs/src/mlpcs.rs:220-243 and pcs/src/ipa.rs:214-298."""
    assert py.compute_pr([0, 0, 0]) == [1]
    assert py.compute_pr([1, 0, 1]) == [0, 0, 0, 0, 0, 1]
    r = [5, 6, 7]
    pr = py.compute_pr(r)
    for x in (2, 3, 10):
        assert py.poly_eval(pr, x) == py.eval_pr(r, x)
    # h = f*rev(g) + rev(f)*g has 2 * <f, g> in the middle: 32 and 14 in the reference's tests
    for a, b, ip in (([1, 2, 3], [4, 5, 6], 32), ([1, 2, 3], [4, 5], 14)):
        L = max(len(a), len(b))
        aa, bb = a + [0] * (L - len(a)), b + [0] * (L - len(b))
        hpoly = py.poly_add(py.poly_mul(aa, bb[::-1]), py.poly_mul(aa[::-1], bb))
        assert hpoly[L - 1] == 2 * ip
        assert py.compute_s_polynomial(a, b) == py.trim(hpoly[L:])
    assert [hx(c) for c in G["s_poly"]["a123_b456"]] == py.compute_s_polynomial([1, 2, 3], [4, 5, 6])


def test_mlpcs_open_evaluation_matches_mle():
    """pcs/src/mlpcs.rs:245-319: proof.evaluation == DenseMultilinearExtension::evaluate (index bit j <-> variable j)."""
    rnd = random.Random(5)
    n = 3
    kz = py.KZG(2 * (1 << n), py.g1_mul(py.G1_GEN, 3), 777)
    poly = [rnd.randrange(FR) for _ in range(1 << n)]
    pt = [rnd.randrange(FR) for _ in range(n)]
    proof = py.mlpcs_open(kz, poly, pt, py.Transcript(b"mlpcs"))
    assert proof["evaluation"] == py.mle_evaluate(poly, pt)


def test_hyperplonk_restatement_golden():
    """proof.rs:63-301 via the Python restatement on the reference's Fibonacci circuits; also checks that the witness
    really satisfies the circuit and that a broken copy constraint is caught (transition_circuit.rs:153-205)."""
    from oracle import fastkzg

    fastkzg.install_fast_s_polynomial()
    k = G["kzg_test"]
    gen = (hx(k["g"][0]), hx(k["g"][1]))
    okzg = fastkzg.FastKZG(64, gen, hx(k["tau"]))
    c1, w1 = py.fibonacci_circuit_and_trace()
    c2, w2 = py.modified_fibonacci_circuit_and_trace()
    assert c1.num_cols() == 4 and c2.num_cols() == 8 and len(c1.public_values()) == 2
    ids, perm = c1.permutation()
    assert sorted(perm) == ids == list(range(1, 33))  # a permutation of 1..32, no zeros (circuit.rs warning)
    h1 = py.hyperplonk_prove([c1], [w1], okzg)
    g = G["hyperplonk"]["fibonacci"]
    assert h1["state_end"] == g["state_end"]
    assert py.ser_g1(h1["witness_commitment"][0]).hex() == g["witness_commitment"]
    assert h1["trace_proofs"][0]["zc_polys"][0] == [hx(c) for c in g["zc_round0"]]
    h2 = py.hyperplonk_prove([c1, c2], [w1, w2], okzg)
    assert h2["state_end"] == G["hyperplonk"]["multitrace"]["state_end"]
    # every zero-check round polynomial of a satisfied circuit sums to the running claim starting from 0
    tp = h1["trace_proofs"][0]
    assert (py.poly_eval(tp["zc_polys"][0], 0) + py.poly_eval(tp["zc_polys"][0], 1)) % FR == 0
    bad = [list(col) for col in w1]
    bad[1][3] = 99
    with pytest.raises(AssertionError):
        c1.check_constraints(bad)
