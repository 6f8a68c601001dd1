"""Generate tests/golden/golden.json from oracle/pyref.py (Python ints + the `blake3` wheel, which wraps the same
upstream Rust crate the reference uses).  The reference itself (Rust/arkworks) cannot run in this image, so these
vectors pin the RESTATEMENT, not the arkworks binary ("parity unpinned", see oracle/README.md).  Shapes follow the
reference's own tests: sumcheck.rs:159-230, zerocheck.rs:85-211, kzg.rs:119-151, eq_eval.rs:53-75, mlpcs.rs:220-243,
ipa.rs:214-298.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyref as py  # noqa: E402

FR = py.FR
H = lambda v: "%064x" % v  # noqa: E731
out = {}

# ---- transcript ----------------------------------------------------------------------------------------------------------
tr = {}
for dom in [b"", b"sumcheck_test", b"zerocheck_test", b"hyperplonk_proof", b"x" * 100]:
    t = py.Transcript(dom)
    s0 = t.state.hex()
    f1 = t.draw_field_element()
    t.append_usize(3)
    t.append_fr(48)
    t.append_fr_vec([0, 38, 10])
    f2 = t.draw_field_element()
    c = t.draw_challenge(17).hex()
    tr[dom.decode()] = dict(state0=s0, fe1=H(f1), fe2=H(f2), challenge17=c, state_end=t.state.hex())
out["transcript"] = tr

# ---- sumcheck: the reference's test_sumcheck_proof (sumcheck.rs:159-230) -----------------------------------------------------
g1 = [((i >> 0) & 1) + 2 * ((i >> 1) & 1) + 3 * ((i >> 2) & 1) for i in range(8)]
g2 = [((i >> 0) & 1) * 2 * ((i >> 1) & 1) + 3 * ((i >> 0) & 1) * ((i >> 2) & 1) for i in range(8)]
h = py.e_mul(py.e_in(0), py.e_in(1))
cs = sum(a * b for a, b in zip(g1, g2)) % FR
t = py.Transcript(b"sumcheck_test")
rp, pt, ev = py.sumcheck_prove(3, [g1, g2], h, cs, t)
out["sumcheck_test"] = dict(claimed_sum=H(cs), r_polys=[[H(c) for c in p] for p in rp], point=[H(x) for x in pt],
                            evaluation=H(ev), state_end=t.state.hex())

# ---- zerocheck: test_zerocheck_proof and test_zerocheck_proof_not_zero (zerocheck.rs:85-211) ------------------------------------
for name, g2v in [("zerocheck_test", [i * i for i in range(8)]), ("zerocheck_test_not_zero", [0, 1, 4, 9, 16, 25, 36, 50])]:
    g1v = list(range(8))
    hz = py.e_sub(py.e_mul(py.e_in(0), py.e_in(0)), py.e_in(1))
    t = py.Transcript(b"zerocheck_test")
    rp, pt, ev, z = py.zerocheck_prove(3, [g1v, g2v], hz, t)
    out[name] = dict(g2=g2v, r_polys=[[H(c) for c in p] for p in rp], point=[H(x) for x in pt], evaluation=H(ev),
                     z=[H(x) for x in z], state_end=t.state.hex())

# ---- a seeded degree-3 product and a mixed expression over 6 variables -----------------------------------------------------------
rnd = random.Random(20261018)
n = 6
tabs = [[rnd.randrange(FR) for _ in range(1 << n)] for _ in range(4)]
h3 = py.e_mul(py.e_mul(py.e_in(0), py.e_in(1)), py.e_in(2))
cs3 = sum(a * b * c for a, b, c in zip(*tabs[:3])) % FR
t = py.Transcript(b"sumcheck_bench")
rp, pt, ev = py.sumcheck_prove(n, tabs, h3, cs3, t)
out["product3_n6"] = dict(tables=[[H(x) for x in tb] for tb in tabs], claimed_sum=H(cs3),
                          r_polys=[[H(c) for c in p] for p in rp], point=[H(x) for x in pt], evaluation=H(ev),
                          state_end=t.state.hex())
hm = py.e_add(py.e_sub(py.e_mul(py.e_in(0), py.e_in(1)), py.e_in(3)), py.e_mul(py.e_const(7), py.e_mul(py.e_in(2), py.e_in(2))))
t = py.Transcript(b"mixed")
rp, pt, ev = py.sumcheck_prove(n, tabs, hm, 123, t)  # a false claim: the prover still runs (zerocheck.rs:161-211)
out["mixed_n6"] = dict(r_polys=[[H(c) for c in p] for p in rp], point=[H(x) for x in pt], evaluation=H(ev),
                       state_end=t.state.hex())

# ---- eq table (eq_eval.rs:53-75) --------------------------------------------------------------------------------------------------
pt5 = [rnd.randrange(FR) for _ in range(5)]
out["eq_n5"] = dict(point=[H(x) for x in pt5], table=[H(x) for x in py.fast_eq_eval_hypercube(5, pt5)])

# ---- KZG: the reference's test_kzg shape (kzg.rs:119-151) with a fixed generator and tau -------------------------------------------
g = py.g1_mul(py.G1_GEN, 7)
tau = 0x1234567890ABCDEF1234567890ABCDEF
kzg = py.KZG(4, g, tau)
poly = [2, 1, 3]
com = kzg.commit(poly)
x, y, pr = kzg.open(poly, 5)
out["kzg_test"] = dict(g=[H(g[0]), H(g[1])], tau=H(tau), srs=[[H(p[0]), H(p[1])] for p in kzg.g1_points],
                       commitment=[H(com[0]), H(com[1])], commitment_bytes=py.ser_g1(com).hex(), y=H(y),
                       proof=[H(pr[0]), H(pr[1])], identity_bytes=py.ser_g1(None).hex())
# a 64-point MSM with seeded scalars (incl. 0, 1, r-1) on the same SRS shape
kz64 = py.KZG(63, g, tau)
sc = [0, 1, FR - 1] + [rnd.randrange(FR) for _ in range(61)]
m64 = py.msm_naive(kz64.g1_points, sc)
out["msm64"] = dict(scalars=[H(s) for s in sc], result=[H(m64[0]), H(m64[1])])

# ---- P_r and S polynomial (mlpcs.rs:220-243, ipa.rs:214-298) ---------------------------------------------------------------------------
out["pr"] = dict(r000=[H(c) for c in py.compute_pr([0, 0, 0])], r101=[H(c) for c in py.compute_pr([1, 0, 1])])
out["s_poly"] = dict(a123_b456=[H(c) for c in py.compute_s_polynomial([1, 2, 3], [4, 5, 6])],
                     a123_b45=[H(c) for c in py.compute_s_polynomial([1, 2, 3], [4, 5])])

# ---- HyperPlonk driver on the reference's integration-test circuits (hyperplonk/tests/test_basic_proof.rs:17-196) -----------------
from oracle import fastkzg  # noqa: E402  (C++ oracle for the MSMs; same sums)

fastkzg.install_fast_s_polynomial()
c1, w1 = py.fibonacci_circuit_and_trace()
c2, w2 = py.modified_fibonacci_circuit_and_trace()
okzg = fastkzg.FastKZG(64, g, tau)
hp1 = py.hyperplonk_prove([c1], [w1], okzg)
hp2 = py.hyperplonk_prove([c1, c2], [w1, w2], okzg)
out["hyperplonk"] = dict(
    fibonacci=dict(state_end=hp1["state_end"], witness_commitment=py.ser_g1(hp1["witness_commitment"][0]).hex(),
                   zc_round0=[H(c) for c in hp1["trace_proofs"][0]["zc_polys"][0]],
                   perm_point=[H(x) for x in hp1["trace_proofs"][0]["perm_point"]]),
    multitrace=dict(state_end=hp2["state_end"]))

# ---- MLEvalProof::prove (pcs/src/mlpcs.rs:83-124): 5 variables, seeded (drawn last so earlier vectors keep their values) ----------
poly5 = [rnd.randrange(FR) for _ in range(32)]
point5 = [rnd.randrange(FR) for _ in range(5)]
t = py.Transcript(b"mlpcs_golden")
com5 = okzg.commit(poly5)
pf5 = py.mlpcs_open(okzg, poly5, point5, t)
op = lambda o: dict(x=H(o[0]), y=H(o[1]), proof_bytes=py.ser_g1(o[2]).hex())  # noqa: E731
out["mlpcs_n5"] = dict(poly=[H(x) for x in poly5], point=[H(x) for x in point5], commitment_bytes=py.ser_g1(com5).hex(),
                       evaluation=H(pf5["evaluation"]), s_comm_bytes=py.ser_g1(pf5["s_comm"]).hex(),
                       poly_opening=op(pf5["poly_opening"]), poly_opening_inv=op(pf5["poly_opening_inv"]),
                       s_opening=op(pf5["s_opening"]), s_opening_inv=op(pf5["s_opening_inv"]), state_end=t.state.hex())

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path, os.path.getsize(path), "bytes")
