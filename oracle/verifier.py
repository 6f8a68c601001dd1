"""CPU ORACLE (test infrastructure, NOT product code) -- the reference's VERIFIER side.

PARITY UNPINNED (see oracle/README.md).  Restates, on Python ints over oracle/pyref.py with the pairing from the C++
oracle (oracle/pairing.hpp), every `verify` of the reference so that the tests can do what the reference's own tests do:
prove -> verify -> accept, tamper -> reject.  Proofs are the dicts oracle/pyref.py's provers
return; tests/util.py converts the product's proof objects to the same shape.

  kzg_verify          pcs/src/kzg.rs:98-108
  mlpcs_verify        pcs/src/mlpcs.rs:126-161
  zerocheck_verify    hyperplonk/src/piops/zerocheck.rs:51-75
  multiset_verify     hyperplonk/src/piops/multiset_check.rs:184-290
  permutation_verify  hyperplonk/src/piops/permutation_check.rs:61-92
  hyperplonk_verify   hyperplonk/src/proof/proof.rs:63-122 (verifying key), 304-523
(the sumcheck verifier, sumcheck.rs:116-150, is pyref.sumcheck_verify)
"""
from __future__ import annotations

from . import coracle as co
from . import pyref as py

FR = py.FR


class VerifierKey:
    """The verifier's half of `KZG::trusted_setup` (kzg.rs:35-59): g1, g2 and tau*g2.  The reference draws both
    generators from the RNG; here g2 = g2_scalar * (the standard BN254 G2 generator)."""

    def __init__(self, g1, tau, g2_scalar=1):
        self.g1 = g1
        self.g2 = co.g2_mul(co.g2_generator(), co.fr1(g2_scalar))
        self.tau_g2 = co.g2_mul(self.g2, co.fr1(tau))  # kzg.rs:52


def kzg_verify(vk: VerifierKey, commitment, opening) -> bool:
    """kzg.rs:98-108: e(C - y g1, g2) == e(proof, tau g2 - x g2), checked as e(C - y g1, g2) e(-proof, tau g2 - x g2) == 1."""
    x, y, proof = opening
    lhs_g1 = py.g1_add(commitment, py.g1_neg(py.g1_mul(vk.g1, y % FR)))
    rhs_g2 = co.g2_add(vk.tau_g2, co.g2_neg(co.g2_mul(vk.g2, co.fr1(x))))
    ok, _ = co.pairing_product([co.g1_to_bytes(lhs_g1), co.g1_to_bytes(py.g1_neg(proof))], [vk.g2, rhs_g2])
    return ok


def mlpcs_verify(vk: VerifierKey, commitment, proof: dict, tr: py.Transcript) -> bool:
    """mlpcs.rs:126-161"""
    tr.append_fr_vec(proof["evaluation_point"])
    tr.append_fr(proof["evaluation"])
    tr.append_g1(proof["s_comm"])
    r = tr.draw_field_element()
    r_inv = py.fr_inv(r)
    # the reference does not compare the openings' x with r / r^-1 (mlpcs.rs:144-147); neither does this restatement
    ok = [kzg_verify(vk, commitment, proof["poly_opening"]), kzg_verify(vk, commitment, proof["poly_opening_inv"]),
          kzg_verify(vk, proof["s_comm"], proof["s_opening"]), kzg_verify(vk, proof["s_comm"], proof["s_opening_inv"])]
    if not all(ok):
        return False
    pr_r = py.eval_pr(proof["evaluation_point"], r)
    pr_r_inv = py.eval_pr(proof["evaluation_point"], r_inv)
    lhs = (proof["poly_opening"][1] * pr_r_inv + proof["poly_opening_inv"][1] * pr_r) % FR
    rhs = (r * proof["s_opening"][1] + r_inv * proof["s_opening_inv"][1] + 2 * proof["evaluation"]) % FR
    return lhs == rhs


def zerocheck_verify(num_vars, r_polys, tr: py.Transcript, claimed_sum=0, sumcheck_num_vars=None):
    """zerocheck.rs:51-75.  Returns (point, evaluation); raises ValueError like the reference's Err."""
    z = [tr.draw_field_element() for _ in range(num_vars)]
    if claimed_sum % FR != 0:
        raise ValueError("Sumcheck claimed sum is not zero")
    if (num_vars if sumcheck_num_vars is None else sumcheck_num_vars) != num_vars:
        raise ValueError("Sumcheck proof num_vars does not match zerocheck num_vars")
    point, ev = py.sumcheck_verify(num_vars, 0, r_polys, tr)
    return point, ev * py.fr_inv(py.eq_eval(z, point)) % FR


def multiset_verify(proof: dict, num_vars, tr: py.Transcript, vk: VerifierKey, left_h_eval, right_h_eval,
                    multiplicities_eval=None):
    """multiset_check.rs:184-290.  `*_eval` are (point, evaluation) claims verified by the caller; Subset mode iff
    `multiplicities_eval` is given.  Raises ValueError like the reference's Err."""
    gamma = tr.draw_field_element()
    tr.append_g1(proof["denom_left_commitment"])
    tr.append_g1(proof["denom_right_commitment"])
    lam = tr.draw_field_element()
    alpha = tr.draw_field_element()
    z = [tr.draw_field_element() for _ in range(len(left_h_eval[0]))]
    if proof.get("claimed_sum", 0) % FR != 0:
        raise ValueError("Multiset equality sumcheck claimed sum is not zero")
    point, ev = py.sumcheck_verify(num_vars, 0, proof["r_polys"], tr)
    ok_l = mlpcs_verify(vk, proof["denom_left_commitment"], proof["opening_proof_denom_left"], tr)
    ok_r = mlpcs_verify(vk, proof["denom_right_commitment"], proof["opening_proof_denom_right"], tr)
    if not ok_l or not ok_r:
        raise ValueError("Multiset equality opening proof verification failed")
    if proof["opening_proof_denom_left"]["evaluation_point"] != point or \
            proof["opening_proof_denom_right"]["evaluation_point"] != point:
        raise ValueError("Multiset equality opening proof evaluation point does not match sumcheck")
    if list(left_h_eval[0]) != point or list(right_h_eval[0]) != point:
        raise ValueError("Multiset equality h evaluation point does not match sumcheck")
    m = 1
    if multiplicities_eval is not None:
        if list(multiplicities_eval[0]) != point:
            raise ValueError("Multiset equality multiplicities evaluation point does not match sumcheck")
        m = multiplicities_eval[1]
    dl, dr = proof["opening_proof_denom_left"]["evaluation"], proof["opening_proof_denom_right"]["evaluation"]
    zc = (dl * (gamma + left_h_eval[1]) - 1 + lam * (dr * (gamma + right_h_eval[1]) - m)) % FR
    final = (zc * py.eq_eval(z, point) * alpha + dl - dr) % FR
    if final != ev:
        raise ValueError("Multiset equality final evaluation does not match sumcheck")


def permutation_verify(proof: dict, num_vars, tr: py.Transcript, vk: VerifierKey, left_h_eval, right_h_eval, id_eval,
                       perm_eval):
    """permutation_check.rs:61-92"""
    alpha = tr.draw_field_element()
    left_hat = (left_h_eval[0], (id_eval[1] + alpha * left_h_eval[1]) % FR)
    right_hat = (right_h_eval[0], (perm_eval[1] + alpha * right_h_eval[1]) % FR)
    multiset_verify(proof, num_vars, tr, vk, left_hat, right_hat)


def hyperplonk_vk(circuits, kzg: py.KZG):
    """proof.rs:63-122: per trace, commitments to the public columns, the id and the permutation polynomial."""
    vks = []
    for c in circuits:
        ids, perm = c.permutation()
        vks.append(dict(circuit=c, public_columns_commitments=[kzg.commit(p) for p in c.public_values()],
                        id_commitment=kzg.commit(ids), permutation_commitment=kzg.commit(perm)))
    return vks


def _claim(opening: dict):
    return opening["evaluation_point"], opening["evaluation"]


def _verify_opening(vk, comm, opening, expected_point, expected_num_vars, tr) -> bool:
    """proof.rs:305-326"""
    if len(opening["evaluation_point"]) != expected_num_vars:
        return False
    if expected_point is not None and opening["evaluation_point"] != list(expected_point):
        return False
    return mlpcs_verify(vk, comm, opening, tr)


def hyperplonk_verify(proof: dict, trace_vks, vk: VerifierKey):
    """proof.rs:493-523 with verify_trace_proof (:400-491).  Raises ValueError like the reference's Err."""
    tr = py.Transcript(b"hyperplonk_proof")
    for com in proof["witness_commitment"]:
        tr.append_g1(com)
    if len(trace_vks) != len(proof["trace_proofs"]):
        raise ValueError("Number of trace VKS and proofs mismatch")
    for wcom, tvk, tp in zip(proof["witness_commitment"], trace_vks, proof["trace_proofs"]):
        c = tvk["circuit"]
        alpha = tr.draw_field_element()
        log2_cols, log2_rows = c.num_cols().bit_length() - 1, c.num_rows.bit_length() - 1
        zc_point, zc_eval = zerocheck_verify(len(tp["zc_polys"]), tp["zc_polys"], tr)
        if len(zc_point) != log2_rows:
            raise ValueError("Zero check evaluation claim point length mismatch")
        trace_claim = _claim(tp["opening_permutation_trace"])
        permutation_verify(tp["permutation"], len(tp["permutation"]["r_polys"]), tr, vk, trace_claim, trace_claim,
                           _claim(tp["opening_id"]), _claim(tp["opening_permutation"]))
        col_evals = []  # get_and_verify_column_evaluations (:331-382)
        for col, opening in enumerate(tp["openings_zero_check"]):
            want = list(zc_point) + [(col >> i) & 1 for i in range(log2_cols)]
            if opening["evaluation_point"] != want:
                raise ValueError("Zero check opening point mismatch")
            if not mlpcs_verify(vk, wcom, opening, tr):
                raise ValueError("Zero check opening verification failed")
            col_evals.append(opening["evaluation"])
        for i, opening in enumerate(tp["openings_public"]):
            if not _verify_opening(vk, tvk["public_columns_commitments"][i], opening, zc_point, log2_rows, tr):
                raise ValueError("Public opening verification failed")
            col_evals.append(opening["evaluation"])
        recomputed = 0  # recover_zerocheck_expr_evaluation (:384-398)
        for i, e in enumerate(c.zero_check_expressions()):
            recomputed = (recomputed + pow(alpha, i, FR) * py.expr_eval_point(e, col_evals)) % FR
        if recomputed != zc_eval:
            raise ValueError("Zero check evaluation mismatch")
        nv = log2_rows + log2_cols
        if not _verify_opening(vk, tvk["id_commitment"], tp["opening_id"], None, nv, tr):
            raise ValueError("ID commitment opening verification failed")
        if not _verify_opening(vk, tvk["permutation_commitment"], tp["opening_permutation"], None, nv, tr):
            raise ValueError("Permutation commitment opening verification failed")
        if not _verify_opening(vk, wcom, tp["opening_permutation_trace"], None, nv, tr):
            raise ValueError("Permutation trace commitment opening verification failed")
    return tr.state.hex()
