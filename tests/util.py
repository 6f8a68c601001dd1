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
