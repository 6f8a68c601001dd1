"""GPU parity: KZG commit (MSM), KZG open and the SRS generator against the oracle, bit-exact; edge cases the
reference accepts (empty input, truncation, zero scalars, identity result, repeated / opposite points, infinity
in the SRS); BASELINE.json sizes through the closed form commit(p) = p(tau) * g."""
import json
import os
import random

import numpy as np
import pytest

import quill_zkvm_b200 as q
from oracle import coracle as co
from oracle import pyref as py
from tests import util

pytestmark = pytest.mark.gpu
FR = py.FR
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
hx = lambda s: int(s, 16)  # noqa: E731
GEN = py.g1_mul(py.G1_GEN, 7)
TAU = 0x1234567890ABCDEF1234567890ABCDEF


@pytest.fixture(scope="module")
def kzg_big(ctx):
    k = q.KZG.trusted_setup(ctx, (1 << 20) - 1, co.g1_to_bytes(GEN), co.fr1(TAU))
    yield k
    k.srs.free()


def test_reference_kzg_test_golden(ctx):
    g = G["kzg_test"]
    kzg = q.KZG.trusted_setup(ctx, 4, co.g1_to_bytes(GEN), co.fr1(TAU))
    assert kzg.max_degree == 4
    pts = kzg.srs.download()
    assert [co.g1_from_bytes(pts[i]) for i in range(5)] == [(hx(p[0]), hx(p[1])) for p in g["srs"]]
    poly = co.to_mont([2, 1, 3])  # kzg.rs:128
    com = kzg.commit(poly)
    assert co.g1_from_bytes(com) == (hx(g["commitment"][0]), hx(g["commitment"][1]))
    assert ctx.g1_serialize(com).hex() == g["commitment_bytes"]
    pr = kzg.open(poly, co.fr1(5))
    assert co.from_mont(pr.y)[0] == 82 == hx(g["y"])
    assert co.g1_from_bytes(pr.proof) == (hx(g["proof"][0]), hx(g["proof"][1]))
    with pytest.raises(AssertionError):  # kzg.rs:62-65
        kzg.commit(co.to_mont([1] * 6))
    kzg.srs.free()


def test_msm64_golden_and_uploaded_srs(ctx):
    g = G["msm64"]
    srs = co.srs_generate(co.g1_to_bytes(GEN), co.fr1(TAU), 64, threads=2)
    kzg = q.KZG.from_points(ctx, srs)
    sc = co.to_mont([hx(s) for s in g["scalars"]])
    assert co.g1_from_bytes(kzg.commit(sc)) == (hx(g["result"][0]), hx(g["result"][1]))
    kzg.srs.free()


def test_edge_cases(ctx):
    srs = co.srs_generate(co.g1_to_bytes(GEN), co.fr1(TAU), 40, threads=2)
    kzg = q.KZG.from_points(ctx, srs)
    ident = np.zeros(64, dtype=np.uint8)
    assert np.array_equal(kzg.commit(np.zeros((0, 32), np.uint8)), ident)          # commit(&[]) = identity
    assert ctx.g1_serialize(ident) == bytes(63) + b"\x40"
    assert np.array_equal(kzg.commit(co.to_mont([0] * 17)), ident)                  # all-zero scalars
    sc = util.rand_fr(64, 5)
    assert np.array_equal(kzg.msm_unchecked(sc), co.msm(srs, sc[:40]))              # msm_unchecked truncates
    assert np.array_equal(kzg.commit(sc[:7]), co.msm(srs[:7], sc[:7]))
    for v in ([1], [FR - 1], [1, 1, 1], [2, 1, 3], [FR - 1] * 40, [1 << 253] * 3):
        s = co.to_mont(v)
        assert np.array_equal(kzg.commit(s), co.msm(srs, s)), v
    kzg.srs.free()
    # repeated and opposite points: doubling / cancellation inside buckets and in the reduction
    p = srs[3]
    neg = co.g1_to_bytes(py.g1_neg(co.g1_from_bytes(p)))
    pts = np.stack([p] * 50 + [neg] * 50 + [srs[4]] * 33)
    kz = q.KZG.from_points(ctx, pts)
    for seed in range(3):
        s = np.concatenate([co.to_mont([3] * 50), co.to_mont([3] * 50), util.rand_fr(33, seed)])
        assert np.array_equal(kz.commit(s), co.msm(pts, s))
    s = co.to_mont([5] * 50 + [5] * 50 + [0] * 33)
    assert np.array_equal(kz.commit(s), ident)
    kz.srs.free()
    # the point at infinity inside the SRS (tau = 0 makes every power but the first the identity)
    pts = np.stack([srs[0], np.zeros(64, np.uint8), srs[2], np.zeros(64, np.uint8)])
    kz = q.KZG.from_points(ctx, pts)
    s = util.rand_fr(4, 9)
    assert np.array_equal(kz.commit(s), co.msm(pts, s))
    kz.srs.free()


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 100, 1000, 4097, 1 << 14, 1 << 16])
def test_msm_sizes_vs_oracle(ctx, kzg_big, n):
    """random scalars, SRS prefix of the device-generated powers; oracle = Pippenger on the host"""
    bases = kzg_big.srs.download(0, n)
    sc = util.rand_fr(n, 1000 + n)
    got = kzg_big.commit(sc)
    assert np.array_equal(got, co.msm(bases, sc, mode=1, threads=os.cpu_count() or 1))
    assert co.g1_on_curve(got)


def test_srs_generator_spot_checks(ctx, kzg_big):
    n = len(kzg_big.srs)
    g = co.g1_to_bytes(GEN)
    for i in [0, 1, 2, 7, 8, 9, 255, 256, 65535, 65536, n - 1]:
        want = co.g1_mul(g, co.fr1(pow(TAU, i, FR)))
        assert np.array_equal(kzg_big.srs.download(i, 1)[0], want), i


def test_small_and_skewed_scalars(ctx, kzg_big):
    """witness-like inputs: tiny values and long runs of equal scalars put most points in a few buckets"""
    n = 5000
    bases = kzg_big.srs.download(0, n)
    rnd = random.Random(3)
    for vals in ([rnd.randrange(2) for _ in range(n)], [rnd.randrange(256) for _ in range(n)], [1] * n,
                 [FR - 1] * n, [rnd.choice([0, 1, FR - 1, 1 << 128]) for _ in range(n)]):
        s = co.to_mont(vals)
        assert np.array_equal(kzg_big.commit(s), co.msm(bases, s, mode=1, threads=os.cpu_count() or 1))


def test_msm_2_20_closed_form(ctx, kzg_big):
    """BASELINE.json config 2: 2^20 random scalars.  commit(p) = p(tau) * g, with p(tau) from the oracle's Horner."""
    n = 1 << 20
    buf = ctx.random_fr(n, 2020)
    sc = buf.download().reshape(-1, 32)
    got_dev = kzg_big.commit(buf)
    got_host = kzg_big.commit(sc)
    buf.free()
    y, _ = co.kzg_open_quotient(sc, co.fr1(TAU))
    want = co.g1_mul(co.g1_to_bytes(GEN), y)
    assert np.array_equal(got_dev, want) and np.array_equal(got_host, want)
    # linearity: commit(a) + commit(b) == commit(a + b)
    b2 = util.rand_fr(n, 7)
    s = co.field_op(0, 0, sc, b2)
    assert np.array_equal(co.g1_add(got_dev, kzg_big.commit(b2)), kzg_big.commit(s))


def test_kzg_open_vs_oracle(ctx, kzg_big):
    for n in (1, 2, 3, 63, 64, 65, 4096, 4097, 100000):
        poly = util.rand_fr(n, 40 + n)
        x = util.rand_fr(1, n)[0]
        pr = kzg_big.open(poly, x)
        y, qpoly = co.kzg_open_quotient(poly, x)
        assert np.array_equal(pr.y, y), n
        bases = kzg_big.srs.download(0, max(n - 1, 1))
        assert np.array_equal(pr.proof, co.msm(bases, qpoly, mode=1, threads=os.cpu_count() or 1)), n
    # constant polynomial: quotient is zero, proof is the identity (reachable from mlpcs.rs:321-393)
    pr = kzg_big.open(co.to_mont([9]), co.fr1(4))
    assert co.from_mont(pr.y)[0] == 9 and not pr.proof.any()
    # trailing zero coefficients do not change anything (DensePolynomial trims them)
    a = kzg_big.open(co.to_mont([1, 2, 3, 0, 0]), co.fr1(11))
    b = kzg_big.open(co.to_mont([1, 2, 3]), co.fr1(11))
    assert np.array_equal(a.y, b.y) and np.array_equal(a.proof, b.proof)


def test_msm_2_24_closed_form(ctx):
    """the size BASELINE.json's metric is quoted on"""
    n = 1 << 24
    kzg = q.KZG.trusted_setup(ctx, n - 1, co.g1_to_bytes(GEN), co.fr1(TAU))
    buf = ctx.random_fr(n, 2424)
    sc = buf.download().reshape(-1, 32)
    got = kzg.commit(buf)
    buf.free()
    kzg.srs.free()
    y, _ = co.kzg_open_quotient(sc, co.fr1(TAU))
    assert np.array_equal(got, co.g1_mul(co.g1_to_bytes(GEN), y))


@pytest.mark.parametrize("c", [4, 8, 13, 16, 0])
def test_precomputed_windows_small(ctx, c):
    """qz_srs_precompute: every window shares one bucket set; results must not change"""
    n = 3000
    srs = co.srs_generate(co.g1_to_bytes(GEN), co.fr1(TAU), n, threads=4)
    srs[7] = 0  # a point at infinity inside the SRS
    kz = q.KZG.from_points(ctx, srs).precompute(c)
    rnd = random.Random(c)
    cases = [util.rand_fr(n, 3 + c), co.to_mont([rnd.randrange(4) for _ in range(n)]), co.to_mont([FR - 1] * n),
             util.rand_fr(17, 1), np.zeros((0, 32), np.uint8), co.to_mont([1])]
    for s in cases:
        assert np.array_equal(kz.commit(s), co.msm(srs, s, mode=1, threads=os.cpu_count() or 1))
    pr = kz.open(cases[0], co.fr1(12345))
    y, qpoly = co.kzg_open_quotient(cases[0], co.fr1(12345))
    assert np.array_equal(pr.y, y) and np.array_equal(pr.proof, co.msm(srs, qpoly, mode=1, threads=4))
    kz.srs.free()


def test_precomputed_windows_2_20_and_2_24(ctx):
    for logn in (20, 24):
        n = 1 << logn
        kzg = q.KZG.trusted_setup(ctx, n - 1, co.g1_to_bytes(GEN), co.fr1(TAU)).precompute()
        buf = ctx.random_fr(n, 31 + logn)
        sc = buf.download().reshape(-1, 32)
        got = kzg.commit(buf)
        half = kzg.commit(sc[: n // 2 + 3])  # a shorter polynomial on the same table
        buf.free()
        kzg.srs.free()
        y, _ = co.kzg_open_quotient(sc, co.fr1(TAU))
        assert np.array_equal(got, co.g1_mul(co.g1_to_bytes(GEN), y)), logn
        y2, _ = co.kzg_open_quotient(sc[: n // 2 + 3], co.fr1(TAU))
        assert np.array_equal(half, co.g1_mul(co.g1_to_bytes(GEN), y2)), logn


def test_msm_2_26_closed_form(ctx):
    """the upper end of north_star's range (2^16..2^26): 4 GiB of bases, 2 GiB of scalars; commit(p) = p(tau) * g"""
    n = 1 << 26
    kzg = q.KZG.trusted_setup(ctx, n - 1, co.g1_to_bytes(GEN), co.fr1(TAU))
    buf = ctx.random_fr(n, 2626)
    got = kzg.commit(buf)
    sc = buf.download().reshape(-1, 32)
    buf.free()
    kzg.srs.free()
    y, _ = co.kzg_open_quotient(sc, co.fr1(TAU))
    assert np.array_equal(got, co.g1_mul(co.g1_to_bytes(GEN), y))


@pytest.mark.parametrize("weights", ["1,1", "2,5,12", "1,1,1,1,1,1,1,1", "1,1000"])
def test_streamed_msm_segments(ctx, kzg_big, weights, monkeypatch):
    """the streamed MSM (point ranges prepared on a second stream, one bucket-set copy per range, merged before the
    reduction) must not change a bit: forced at small sizes, host and device scalars, skewed and ragged inputs"""
    monkeypatch.setenv("QZ_MSM_SEGMENTS", weights)
    monkeypatch.setenv("QZ_MSM_SEGMENTS_DEV", weights)
    rnd = random.Random(len(weights))
    for n in (1, 2, 255, 256, 257, 1000, 5000, 70001):
        bases = kzg_big.srs.download(0, n)
        cases = [util.rand_fr(n, 77 + n)]
        if n == 5000:
            cases += [co.to_mont([rnd.randrange(3) for _ in range(n)]), co.to_mont([FR - 1] * n), co.to_mont([0] * n)]
        for sc in cases:
            want = co.msm(bases, sc, mode=1, threads=os.cpu_count() or 1)
            assert np.array_equal(kzg_big.commit(sc), want), (n, weights)
            dev = ctx.upload(sc)
            assert np.array_equal(kzg_big.commit(dev), want), (n, weights, "device scalars")
            dev.free()
    # KZG::open and the precomputed-window layout on top of the segments
    n = 3000
    srs = co.srs_generate(co.g1_to_bytes(GEN), co.fr1(TAU), n, threads=4)
    kz = q.KZG.from_points(ctx, srs).precompute(13)
    sc = util.rand_fr(n, 5)
    assert np.array_equal(kz.commit(sc), co.msm(srs, sc, mode=1, threads=4))
    pr = kz.open(sc, co.fr1(999))
    y, qpoly = co.kzg_open_quotient(sc, co.fr1(999))
    assert np.array_equal(pr.y, y) and np.array_equal(pr.proof, co.msm(srs, qpoly, mode=1, threads=4))
    kz.srs.free()


def test_streamed_msm_default_host_path_2_20(ctx, kzg_big):
    """host scalars at 2^20 take the streamed path by default (>= 2^19): same commitment as device-resident scalars"""
    n = 1 << 20
    buf = ctx.random_fr(n, 909)
    sc = buf.download().reshape(-1, 32)
    a = kzg_big.commit(buf)
    b = kzg_big.commit(sc)
    buf.free()
    y, _ = co.kzg_open_quotient(sc, co.fr1(TAU))
    assert np.array_equal(a, b) and np.array_equal(a, co.g1_mul(co.g1_to_bytes(GEN), y))


def test_kzg_open_large_closed_form(ctx):
    """KZG::open beyond 2^20 coefficients (the carry scan of the quotient links several level-2 groups per thread, the
    last run ragged): y = p(x) from the oracle's Horner, proof = commit(q) = ((p(tau) - y) / (tau - x)) * g"""
    n = (1 << 22) + 4099
    kzg = q.KZG.trusted_setup(ctx, n - 1, co.g1_to_bytes(GEN), co.fr1(TAU))
    buf = ctx.random_fr(n, 4242)
    sc = buf.download().reshape(-1, 32)
    x = util.rand_fr(1, 99)[0]
    pr = kzg.open(buf, x)
    buf.free()
    kzg.srs.free()
    y, _ = co.kzg_open_quotient(sc, x)
    ptau, _ = co.kzg_open_quotient(sc, co.fr1(TAU))
    yi, pti, xi = (co.from_mont(v.reshape(1, 32))[0] for v in (y, ptau, x))
    qtau = (pti - yi) * pow((TAU - xi) % FR, -1, FR) % FR
    assert np.array_equal(pr.y, y)
    assert np.array_equal(pr.proof, co.g1_mul(co.g1_to_bytes(GEN), co.fr1(qtau)))


def test_commit_split_and_accumulate_stats(ctx, kzg_big):
    """qz_msm_split without a communicator is KZG::commit (one rank takes the whole index range; the N > 1 case runs in
    tools/multi_gpu_check.py and the bench's HyperPlonk leg); qz_msm_accumulate_stats counts the additions the bucket
    accumulation executes (non-zero digits) and times its launches."""
    n = 1 << 14
    sc = util.rand_fr(n, 4242)
    sc[5] = 0  # a zero scalar: every digit zero, no addition
    want = kzg_big.commit(sc)
    assert np.array_equal(kzg_big.commit_split(sc), want)
    d = ctx.upload(sc)
    assert np.array_equal(kzg_big.commit_split(d), want)
    with pytest.raises(AssertionError):
        kzg_big.commit_split(util.rand_fr((1 << 20) + 1, 1))
    ctx.msm_accumulate_stats(1)
    kzg_big.commit(d)
    kzg_big.commit(d)
    ms, adds, launches = ctx.msm_accumulate_stats(-1)
    digits = ctx.last_stat(1)
    assert launches == 2 and ms > 0
    assert (n - 2) * digits * 2 * 0.9 < adds <= (n - 1) * digits * 2  # one scalar is zero; a digit is zero with probability 2^-c
    kzg_big.commit(d)
    assert ctx.msm_accumulate_stats(0)[2] == 0  # collection stopped
    d.free()


@pytest.mark.parametrize("levels", [1, 3, 8])
def test_pair_levels(ctx, kzg_big, levels, monkeypatch):
    """QZ_MSM_PAIR_LEVELS: equal-key neighbours of the sorted list are added in affine coordinates with a shared
    inversion before the XYZZ accumulation (csrc/msm.cu msm_pair_*; CPU model: tests/test_device_models.py).  Not a bit
    may change: ragged sizes, skewed scalars, P + P / P - P / infinity inside buckets, precomputed windows, streamed
    ranges, KZG::open, the 2^20 closed form."""
    monkeypatch.setenv("QZ_MSM_PAIR_LEVELS", str(levels))
    threads = os.cpu_count() or 1
    rnd = random.Random(levels)
    for n in (63, 64, 65, 100, 1000, 4097, 1 << 14):
        bases = kzg_big.srs.download(0, n)
        cases = [util.rand_fr(n, 500 + n)]
        if n == 4097:
            cases += [co.to_mont([rnd.randrange(2) for _ in range(n)]), co.to_mont([rnd.randrange(256) for _ in range(n)]),
                      co.to_mont([FR - 1] * n), co.to_mont([rnd.choice([0, 1, FR - 1, 1 << 128]) for _ in range(n)])]
        for sc in cases:
            want = co.msm(bases, sc, mode=1, threads=threads)
            assert np.array_equal(kzg_big.commit(sc), want), n
            dev = ctx.upload(sc)
            assert np.array_equal(kzg_big.commit(dev), want), (n, "device scalars")
            dev.free()
    # repeated and opposite points, infinity in the SRS
    srs = co.srs_generate(co.g1_to_bytes(GEN), co.fr1(TAU), 40, threads=2)
    p = srs[3]
    neg = co.g1_to_bytes(py.g1_neg(co.g1_from_bytes(p)))
    pts = np.stack([p] * 50 + [neg] * 50 + [srs[4]] * 33 + [np.zeros(64, np.uint8)] * 7)
    kz = q.KZG.from_points(ctx, pts)
    for s in (np.concatenate([co.to_mont([3] * 100), util.rand_fr(40, 1)]), co.to_mont([5] * 140), util.rand_fr(140, 2)):
        assert np.array_equal(kz.commit(s), co.msm(pts, s))
    kz.srs.free()
    # precomputed windows (one shared bucket set: long runs), with streamed ranges on top, and KZG::open
    n = 3000
    srs = co.srs_generate(co.g1_to_bytes(GEN), co.fr1(TAU), n, threads=4)
    srs[7] = 0
    for c in (5, 13, 0):
        kz = q.KZG.from_points(ctx, srs).precompute(c)
        for s in (util.rand_fr(n, 3 + c), co.to_mont([rnd.randrange(4) for _ in range(n)]), co.to_mont([FR - 1] * n)):
            assert np.array_equal(kz.commit(s), co.msm(srs, s, mode=1, threads=threads)), c
        monkeypatch.setenv("QZ_MSM_SEGMENTS", "2,5,12")
        s = util.rand_fr(n, 77)
        assert np.array_equal(kz.commit(s), co.msm(srs, s, mode=1, threads=threads)), (c, "segments")
        monkeypatch.delenv("QZ_MSM_SEGMENTS")
        pr = kz.open(s, co.fr1(12345))
        y, qpoly = co.kzg_open_quotient(s, co.fr1(12345))
        assert np.array_equal(pr.y, y) and np.array_equal(pr.proof, co.msm(srs, qpoly, mode=1, threads=threads))
        kz.srs.free()
    # BASELINE config 2 through the closed form, device and host (streamed) scalars
    n = 1 << 20
    buf = ctx.random_fr(n, 2020 + levels)
    sc = buf.download().reshape(-1, 32)
    got_dev = kzg_big.commit(buf)
    got_host = kzg_big.commit(sc)
    buf.free()
    y, _ = co.kzg_open_quotient(sc, co.fr1(TAU))
    want = co.g1_mul(co.g1_to_bytes(GEN), y)
    assert np.array_equal(got_dev, want) and np.array_equal(got_host, want)
