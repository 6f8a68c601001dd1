"""GPU parity: Fr / Fq Montgomery arithmetic and the G1 group law against the oracle (bit-exact)."""
import random

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyref as py
from tests import util

pytestmark = pytest.mark.gpu
FR, FQ = py.FR, py.FQ


def _edge_and_random(mod, n, seed):
    rnd = random.Random(seed)
    vals = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 253) % mod, (1 << 32) - 1, 1 << 32, (1 << 224) - 1]
    vals += [rnd.randrange(mod) for _ in range(n - len(vals))]
    return vals


@pytest.mark.parametrize("field,mod", [(0, FR), (1, FQ)])
def test_field_ops_bit_exact(ctx, field, mod):
    a = _edge_and_random(mod, 4096, 1 + field)
    b = list(reversed(_edge_and_random(mod, 4096, 3 + field)))
    A, B = co.to_mont(a, mod), co.to_mont(b, mod)
    for op in (0, 1, 2):
        got = ctx.field_op(field, op, A, B)
        want = co.field_op(field, op, A, B)
        assert np.array_equal(got, want), f"field {field} op {op}"
    assert co.from_mont(ctx.field_op(field, 2, A, B), mod) == [x * y % mod for x, y in zip(a, b)]
    # to_mont / from_mont round trip; to_mont accepts non-reduced 256-bit inputs
    raw = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in a[:64]), dtype=np.uint8).reshape(-1, 32)
    assert np.array_equal(ctx.field_op(field, 4, raw), A[:64])
    assert np.array_equal(ctx.field_op(field, 5, A[:64]), raw)
    big = np.full((4, 32), 0xFF, dtype=np.uint8)
    assert co.from_mont(ctx.field_op(field, 4, big), mod) == [((1 << 256) - 1) % mod] * 4


@pytest.mark.parametrize("field,mod", [(0, FR), (1, FQ)])
def test_deferred_reduction_accumulator(ctx, field, mod):
    """sums of unreduced 512-bit products reduced once (ff.cuh wide_mul_acc / wide_reduce) equal the sums of
    Montgomery products, including when the accumulator's 17th word is in use"""
    a = _edge_and_random(mod, 2048, 21 + field)
    b = list(reversed(_edge_and_random(mod, 2048, 23 + field)))
    A, B = co.to_mont(a, mod), co.to_mont(b, mod)
    assert co.from_mont(ctx.field_op(field, 6, A, B), mod) == [(x * y + x * x + y * y) % mod for x, y in zip(a, b)]
    assert co.from_mont(ctx.field_op(field, 7, A, B), mod) == [4096 * x * y % mod for x, y in zip(a, b)]
    # dedicated squaring (36 limb products + a separate reduction)
    assert co.from_mont(ctx.field_op(field, 9, A), mod) == [x * x % mod for x in a]
    # fused two-product multiplier (one Montgomery reduction for a*b + c*d), used by the group law's Y3
    assert co.from_mont(ctx.field_op(field, 8, A, B), mod) == [(x * y + (x + y) * (x - y)) % mod for x, y in zip(a, b)]


@pytest.mark.parametrize("field,mod", [(0, FR), (1, FQ)])
def test_field_inverse(ctx, field, mod):
    a = _edge_and_random(mod, 64, 9)
    A = co.to_mont(a, mod)
    got = co.from_mont(ctx.field_op(field, 3, A), mod)
    assert got == [pow(x, mod - 2, mod) for x in a]
    # the single-thread critical-path inverse (binary extended Euclid, ff.cuh fp_inv_serial): same values, 0 -> 0
    a = _edge_and_random(mod, 512, 10) + [2, 3, 4, (mod + 1) // 2, 1 << 128, (1 << 253) % mod, mod - 2]
    got = co.from_mont(ctx.field_op(field, 10, co.to_mont(a, mod)), mod)
    assert got == [pow(x, mod - 2, mod) for x in a]


def test_fixed_challenge_fold(ctx):
    """a0 + r (a1 - a0) through the table of shifted multiples of r in constant memory (ff.cuh fp_mul_fixed, the fold of the
    large sumcheck passes: sumcheck.rs:81-92) against plain modular arithmetic, edge values included"""
    rnd = random.Random(77)
    for r in (0, 1, FR - 1, (1 << 253) % FR, rnd.randrange(FR), rnd.randrange(FR)):
        a0 = _edge_and_random(FR, 2048, 5)
        a1 = list(reversed(_edge_and_random(FR, 2048, 9)))
        got = co.from_mont(ctx.fold(co.fr1(r), co.to_mont(a0), co.to_mont(a1)))
        assert got == [(x + r * (y - x)) % FR for x, y in zip(a0, a1)], hex(r)


def test_random_fr_generator_is_reduced(ctx):
    buf = ctx.random_fr(1000, 42)
    v = buf.download().reshape(-1, 32)
    buf.free()
    ints = [int.from_bytes(v[i].tobytes(), "little") for i in range(1000)]
    assert all(x < FR for x in ints) and len(set(ints)) == 1000


def test_g1_add_complete(ctx):
    rnd = random.Random(4)
    G = py.G1_GEN
    pts = [py.g1_mul(G, rnd.randrange(1, FR)) for _ in range(20)]
    a = pts + [pts[0], pts[1], None, pts[2], None]
    b = pts[1:] + pts[:1] + [pts[0], py.g1_neg(pts[1]), pts[3], None, None]  # P+P, P+(-P), O+P, P+O, O+O
    A = np.stack([co.g1_to_bytes(p) for p in a])
    B = np.stack([co.g1_to_bytes(p) for p in b])
    got = ctx.g1_add(A, B)
    for i in range(len(a)):
        assert co.g1_from_bytes(got[i]) == py.g1_add(a[i], b[i]), i


def test_g1_scalar_mul_and_serialize(ctx):
    rnd = random.Random(8)
    ks = [0, 1, 2, FR - 1, FR - 2] + [rnd.randrange(FR) for _ in range(11)]
    base = py.g1_mul(py.G1_GEN, 77)
    A = np.stack([co.g1_to_bytes(base)] * len(ks))
    got = ctx.g1_mul(A, co.to_mont(ks))
    for i, k in enumerate(ks):
        want = py.g1_mul(base, k)
        assert co.g1_from_bytes(got[i]) == want
        assert ctx.g1_serialize(got[i]) == py.ser_g1(want)


def test_device_transcript_field_element(ctx):
    import quill_zkvm_b200 as q

    t = q.Transcript(b"sumcheck_test", ctx)
    fe = t.draw_field_element()
    assert co.from_mont(fe)[0] == 15494200051891961783909833458049727794161400064843832742164888701625601212558
    st = co.transcript_new(b"sumcheck_test")
    co.transcript_draw_fr(st)
    assert t.state.tobytes() == st.tobytes()
    t.append_fr(co.fr1(48))
    co.transcript_append(st, py.ser_fr(48))
    p = py.g1_mul(py.G1_GEN, 5)
    t.append_g1(co.g1_to_bytes(p))
    co.transcript_append(st, py.ser_g1(p))
    assert t.state.tobytes() == st.tobytes()
