"""GPU parity: logup denominators 1/(gamma + h(row)) [* m(row)] (hyperplonk/src/piops/multiset_check.rs:43-95)
against the oracle's per-row inversion, bit-exact, in Equality and Subset modes."""
import numpy as np
import pytest

import quill_zkvm_b200 as q
from oracle import coracle as co
from oracle import pyref as py
from tests import util

pytestmark = pytest.mark.gpu
FR = py.FR


def _setup(n, exprs_py, seed):
    tabs = [util.rand_fr(1 << n, seed + t) for t in range(3)]
    store = q.VirtualPolynomialStore(n)
    for t in tabs:
        store.allocate_polynomial(t)
    refs, flat = [], []
    for e in exprs_py:
        nodes, consts = util.expr_from_py(e)
        refs.append(store.new_virtual_from_expr(util.to_qexpr(nodes, consts)))
        flat.append((nodes, consts))
    return tabs, store, refs, flat


@pytest.mark.parametrize("n", [0, 1, 5, 12, 13, 17])
def test_equality_mode(ctx, n):
    h = py.e_add(py.e_in(0), py.e_mul(py.e_const(0x1234), py.e_in(1)))  # id + alpha * h, as permutation_check.rs:13-58 builds it
    tabs, store, (hr,), ((nodes, consts),) = _setup(n, [h], 10 * n)
    gamma = util.rand_fr(1, 99)[0]
    got = q.logup_denominators(ctx, store, hr, gamma)
    want = co.logup_denominators(n, tabs, nodes, [], consts, gamma)
    assert np.array_equal(got, want)
    # definition: out * (gamma + h) == 1
    hv = co.field_op(0, 0, tabs[0], co.field_op(0, 2, np.tile(co.fr1(0x1234), (1 << n, 1)), tabs[1]))
    prod = co.field_op(0, 2, got, co.field_op(0, 0, hv, np.tile(gamma, (1 << n, 1))))
    assert np.array_equal(prod, np.tile(co.fr1(1), (1 << n, 1)))


@pytest.mark.parametrize("n", [3, 12, 14])
def test_subset_mode_with_multiplicities(ctx, n):
    h = py.e_mul(py.e_in(0), py.e_in(1))
    m = py.e_add(py.e_in(2), py.e_const(3))
    tabs, store, (hr, mr), ((nh, ch), (nm, cm)) = _setup(n, [h, m], 7 * n)
    gamma = util.rand_fr(1, 5)[0]
    got = q.logup_denominators(ctx, store, hr, gamma, multiplicities=mr)
    # oracle: one shared consts array, m's Const indices follow h's
    nm2 = [(op, a + (ch.shape[0] if op == 1 else 0), b) for op, a, b in nm]
    want = co.logup_denominators(n, tabs, nh, nm2, np.concatenate([ch, cm]) if ch.shape[0] else cm, gamma)
    assert np.array_equal(got, want)


def test_zero_denominator_panics_like_the_reference(ctx):
    n = 6
    tabs = [util.rand_fr(1 << n, 3)]
    store = q.VirtualPolynomialStore(n)
    store.allocate_polynomial(tabs[0])
    hr = store.new_virtual_from_input(0)
    gamma = co.field_op(0, 1, np.zeros((1, 32), np.uint8), tabs[0][17:18])[0]  # gamma = -table[17]
    with pytest.raises(ZeroDivisionError):
        q.logup_denominators(ctx, store, hr, gamma)
    with pytest.raises(ZeroDivisionError):
        co.logup_denominators(n, tabs, [(0, 0, 0)], [], np.zeros((0, 32), np.uint8), gamma)
