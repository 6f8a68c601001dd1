"""Shared helpers for the tests: seeded inputs and expression shapes (test infrastructure)."""
from __future__ import annotations

import numpy as np

from oracle import coracle as co
from oracle import pyref as py

FR, FQ = py.FR, py.FQ


def rand_fr(n: int, seed: int) -> np.ndarray:
    """n pseudo-random Fr as (n, 32) Montgomery bytes: 252 random bits (< r) taken as the Montgomery form."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return np.ascontiguousarray(a.view(np.uint8).reshape(n, 32))


def small_fr(vals) -> np.ndarray:
    return co.to_mont([v % FR for v in vals])


# expression shapes as node lists (op, a, b): 0 Input, 1 Const, 2 Add, 3 Mul
def expr_product(k: int):
    nodes = [(0, 0, 0)]
    for i in range(1, k):
        nodes.append((0, i, 0))
        nodes.append((3, len(nodes) - 2, len(nodes) - 1))
    return nodes, np.zeros((0, 32), dtype=np.uint8)


def expr_from_py(e):
    """pyref tuple expression -> (nodes, consts Montgomery)."""
    nodes, consts = [], []

    def go(x):
        if x[0] == "in":
            nodes.append((0, x[1], 0))
        elif x[0] == "const":
            consts.append(x[1])
            nodes.append((1, len(consts) - 1, 0))
        else:
            l = go(x[1])
            r = go(x[2])
            nodes.append((2 if x[0] == "add" else 3, l, r))
        return len(nodes) - 1

    go(e)
    return nodes, (co.to_mont(consts) if consts else np.zeros((0, 32), dtype=np.uint8))


def to_qexpr(nodes, consts):
    """node list -> quill_zkvm_b200.VirtualPolyExpr"""
    import quill_zkvm_b200 as q

    built = []
    for op, a, b in nodes:
        if op == 0:
            built.append(q.VirtualPolyExpr.Input(a))
        elif op == 1:
            built.append(q.VirtualPolyExpr.Const(consts[a]))
        elif op == 2:
            built.append(built[a] + built[b])
        else:
            built.append(built[a] * built[b])
    return built[-1]


# ---- product proof objects -> the dict shapes oracle/pyref.py's provers return (input of oracle/verifier.py) ----------
def g1_py(b):
    return co.g1_from_bytes(b)


def kzg_opening_py(o):
    return co.from_mont(o.x)[0], co.from_mont(o.y)[0], g1_py(o.proof)


def opening_py(p) -> dict:
    """quill_zkvm_b200.MLEvalProof -> pyref.mlpcs_open's dict"""
    return dict(evaluation_point=co.from_mont(p.evaluation_point) if len(p.evaluation_point) else [],
                evaluation=co.from_mont(p.evaluation)[0], s_comm=g1_py(p.s_comm),
                poly_opening=kzg_opening_py(p.poly_opening), poly_opening_inv=kzg_opening_py(p.poly_opening_inv),
                s_opening=kzg_opening_py(p.s_opening), s_opening_inv=kzg_opening_py(p.s_opening_inv))


def multiset_py(ms) -> dict:
    return dict(denom_left_commitment=g1_py(ms.denom_left_commitment),
                denom_right_commitment=g1_py(ms.denom_right_commitment),
                claimed_sum=co.from_mont(ms.sumcheck_proof.claimed_sum)[0],
                r_polys=[co.from_mont(p) if len(p) else [] for p in ms.sumcheck_proof.r_polys],
                opening_proof_denom_left=opening_py(ms.opening_proof_denom_left),
                opening_proof_denom_right=opening_py(ms.opening_proof_denom_right))


def hyperplonk_py(proof) -> dict:
    """quill_zkvm_b200.hyperplonk.HyperPlonkProof -> pyref.hyperplonk_prove's dict"""
    tps = []
    for tp in proof.trace_proofs:
        tps.append(dict(
            zc_polys=[co.from_mont(p) if len(p) else [] for p in tp.zero_check_proof.sumcheck_proof.r_polys],
            permutation=multiset_py(tp.permutation_check_proof.multiset_equality_proof),
            openings_zero_check=[opening_py(o) for o in tp.openings_zero_check],
            openings_public=[opening_py(o) for o in tp.openings_public],
            opening_id=opening_py(tp.opening_id), opening_permutation=opening_py(tp.opening_permutation),
            opening_permutation_trace=opening_py(tp.opening_permutation_trace)))
    return dict(witness_commitment=[g1_py(c) for c in proof.witness_commitment], trace_proofs=tps,
                state_end=bytes(proof.transcript_state).hex())


def hyperplonk_vk_py(trace_vks, circuits_py):
    """the product's verifying keys (hyperplonk.HyperPlonk.trace_vks) with the oracle-side circuit descriptions"""
    return [dict(circuit=c, public_columns_commitments=[g1_py(p) for p in v.public_columns_commitments],
                 id_commitment=g1_py(v.id_commitment), permutation_commitment=g1_py(v.permutation_commitment))
            for v, c in zip(trace_vks, circuits_py)]
