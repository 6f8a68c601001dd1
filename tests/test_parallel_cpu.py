"""CPU, world_size 2 over gloo: the sharding the multi-GPU path uses, with the oracle standing in for the
device.  MSM: per-rank partial sums over contiguous index ranges, gathered and added with the group law.  Sumcheck:
tables split by the top variable; per round the ranks exchange partial round polynomials and run the same transcript;
when one entry per rank is left the shards are gathered and the last rounds are replayed by every rank.  Zero-check:
the eq-factored rounds with the weight table sharded like the pairs, then the hand-over in the reference's form."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import coracle as co  # noqa: E402
from oracle import pyref as py  # noqa: E402
from quill_zkvm_b200 import parallel  # noqa: E402

FR = py.FR


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_obj(obj, world):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # bootstrap: the 128-byte id produced on rank 0 reaches every rank
        uid = np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, dtype=np.uint8)
        assert parallel.broadcast_bytes(uid).tolist() == list(range(128))

        # ---- MSM sharding ----
        n = 96
        gen = py.g1_mul(py.G1_GEN, 5)
        srs = co.srs_generate(co.g1_to_bytes(gen), co.fr1(4242), n, threads=1)
        rng = np.random.default_rng(3)
        sc = co.to_mont([int(x) for x in rng.integers(1, 1 << 62, size=n)])
        lo, hi = parallel.shard_range(n, rank, world)
        part = co.msm(srs[lo:hi], sc[lo:hi], mode=1)
        parts = _gather_obj(part.tobytes(), world)
        total = np.zeros(64, dtype=np.uint8)
        for p in parts:
            total = co.g1_add(total, np.frombuffer(p, dtype=np.uint8).copy())
        full = co.msm(srs, sc, mode=1)
        assert total.tobytes() == full.tobytes()

        # ---- sumcheck sharding ----
        nv, k = 5, 3
        import random
        rnd = random.Random(9)
        tabs = [[rnd.randrange(FR) for _ in range(1 << nv)] for _ in range(k)]
        h = py.e_mul(py.e_mul(py.e_in(0), py.e_in(1)), py.e_in(2))
        claimed = sum(a * b * c for a, b, c in zip(*tabs)) % FR
        want = py.sumcheck_prove(nv, tabs, h, claimed, py.Transcript(b"shard"))
        lo, hi = parallel.table_shard_range(nv, rank, world)
        gs = [t[lo:hi] for t in tabs]
        tr = py.Transcript(b"shard")
        tr.append_usize(nv)
        tr.append_fr(claimed)
        polys, point = [], []
        while len(gs[0]) > 1:  # local rounds: every pair (2p, 2p+1) is inside the shard
            msg = []
            for p in range(len(gs[0]) // 2):
                lin = [py.trim([g[2 * p], g[2 * p + 1] - g[2 * p]]) for g in gs]
                msg = py.poly_add(msg, py.expr_eval_poly(h, lin))
            total_msg = []
            for m in _gather_obj(msg, world):  # the per-round exchange: deg+1 field elements per rank
                total_msg = py.poly_add(total_msg, m)
            tr.append_fr_vec(total_msg)
            polys.append(total_msg)
            r = tr.draw_field_element()
            point.append(r)
            gs = [[(g[2 * p] + r * (g[2 * p + 1] - g[2 * p])) % FR for p in range(len(g) // 2)] for g in gs]
        # one entry per rank left: gather (rank order = index order) and finish on every rank
        rest = _gather_obj([g[0] for g in gs], world)
        gs = [[rest[rk][t] for rk in range(world)] for t in range(k)]
        ev = 0
        while len(gs[0]) > 1 or not polys or len(polys) < nv:
            msg = []
            for p in range(len(gs[0]) // 2):
                lin = [py.trim([g[2 * p], g[2 * p + 1] - g[2 * p]]) for g in gs]
                msg = py.poly_add(msg, py.expr_eval_poly(h, lin))
            tr.append_fr_vec(msg)
            polys.append(msg)
            r = tr.draw_field_element()
            point.append(r)
            gs = [[(g[2 * p] + r * (g[2 * p + 1] - g[2 * p])) % FR for p in range(len(g) // 2)] for g in gs]
            if len(gs[0]) == 1:
                ev = py.expr_eval_point(h, [g[0] for g in gs])
                break
        assert (polys, point, ev) == want

        # ---- zero-check sharding, eq-factored rounds (csrc/sumcheck.cu sc_round_zc with G > 1) ----
        # rank g holds pairs [g * N/2G, ...) of the weight table E_1 = eq(., z_1..z_{n-1}); weights fold by addition, which
        # is as local as the pairs; per round the ranks exchange the k+1 evaluations of t_j; with two entries per rank left
        # the eq shard is materialised in the reference's form (last challenge still to fold) and everything is gathered.
        nv, k = 5, 2
        tabs = [[rnd.randrange(FR) for _ in range(1 << nv)] for _ in range(k)]
        hz = py.e_mul(py.e_in(0), py.e_in(1))
        want_z = py.zerocheck_prove(nv, tabs, hz, py.Transcript(b"shard_zc"))
        tr = py.Transcript(b"shard_zc")
        z = [tr.draw_field_element() for _ in range(nv)]
        tr.append_usize(nv)
        tr.append_fr(0)
        lo, hi = parallel.table_shard_range(nv, rank, world)
        gs = [t[lo:hi] for t in tabs]
        e1 = py.fast_eq_eval_hypercube(nv - 1, z[1:])
        weights = e1[lo // 2: hi // 2]
        prefix, prefix_prev, polys, point = 1, 1, [], []

        def lagrange(evals):  # coefficients from evaluations at 0..len-1
            n_pts, coeffs = len(evals), [0] * len(evals)
            for i, yi in enumerate(evals):
                num, den = [1], 1
                for m in range(n_pts):
                    if m != i:
                        num = py.poly_mul(num, [(-m) % FR, 1])
                        den = den * (i - m) % FR
                scale = yi * pow(den, FR - 2, FR) % FR
                num = num + [0] * (n_pts - len(num))
                coeffs = [(c + scale * nc) % FR for c, nc in zip(coeffs, num)]
            return coeffs

        j = 0
        while True:
            pairs = len(gs[0]) // 2
            assert len(weights) == pairs
            part = []
            for x in range(k + 1):
                acc = 0
                for p in range(pairs):
                    term = weights[p]
                    for g in gs:
                        term = term * (g[2 * p] + x * (g[2 * p + 1] - g[2 * p])) % FR
                    acc = (acc + term) % FR
                part.append(acc)
            evals = [sum(col) % FR for col in zip(*_gather_obj(part, world))]  # the per-round exchange
            c = lagrange(evals)
            a, b = prefix * (1 - z[j]) % FR, prefix * (2 * z[j] - 1) % FR
            s_j = py.trim([(a * (c[t] if t <= k else 0) + b * (c[t - 1] if t >= 1 else 0)) % FR for t in range(k + 2)])
            tr.append_fr_vec(s_j)
            polys.append(s_j)
            r = tr.draw_field_element()
            point.append(r)
            prefix_prev, prefix = prefix, prefix * ((r * z[j] + (1 - r) * (1 - z[j])) % FR) % FR
            j += 1
            if len(gs[0]) == 2:
                break  # hand over with the fold by r pending, as the device does at 2^11 entries
            gs = [[(g[2 * p] + r * (g[2 * p + 1] - g[2 * p])) % FR for p in range(pairs)] for g in gs]
            weights = [(weights[2 * p] + weights[2 * p + 1]) % FR for p in range(pairs // 2)]
        eq_local = [prefix_prev * (1 - z[j - 1]) % FR * weights[0] % FR, prefix_prev * z[j - 1] % FR * weights[0] % FR]
        rest = _gather_obj((gs, eq_local), world)  # rank order = index order
        full = [[v for rk in range(world) for v in rest[rk][0][t]] for t in range(k)] + [[v for rk in range(world) for v in rest[rk][1]]]
        h_hat = py.e_mul(hz, py.e_in(k))
        r_pending, ev = point[-1], 0
        while True:  # the reference's rounds on h * eq from here on (sc_tail)
            full = [[(g[2 * p] + r_pending * (g[2 * p + 1] - g[2 * p])) % FR for p in range(len(g) // 2)] for g in full]
            if len(full[0]) == 1:
                ev = py.expr_eval_point(h_hat, [g[0] for g in full])
                break
            msg = []
            for p in range(len(full[0]) // 2):
                lin = [py.trim([g[2 * p], g[2 * p + 1] - g[2 * p]]) for g in full]
                msg = py.poly_add(msg, py.expr_eval_poly(h_hat, lin))
            tr.append_fr_vec(msg)
            polys.append(msg)
            r_pending = tr.draw_field_element()
            point.append(r_pending)
        ev = ev * py.fr_inv(py.eq_eval(z, point)) % FR
        assert (polys, point, ev, z) == (want_z[0], want_z[1], want_z[2] % FR, want_z[3])

        # ---- MLPCS openings dealt to the ranks (hyperplonk.OpeningBatch): the halves before / after the challenge ----
        from quill_zkvm_b200 import hyperplonk as hp
        kz = py.KZG(16, gen, 4242)
        polys_o = [[rnd.randrange(FR) for _ in range(1 << m)] for m in (4, 2, 4, 3, 4)]
        points_o = [[rnd.randrange(FR) for _ in range(m)] for m in (4, 2, 4, 3, 3)]
        seq_tr = py.Transcript(b"deal")
        want_o = [py.mlpcs_open(kz, p, pt, seq_tr) for p, pt in zip(polys_o, points_o)]
        batch = hp.OpeningBatch(None, None, None)
        for p, pt in zip(polys_o, points_o):
            batch.add(np.zeros((len(p), 32), np.uint8), np.zeros((len(pt), 32), np.uint8), None)
        pool = hp.OpeningPool(None, None, nranks=world, rank=rank)
        owner = batch.owners(world, pool)
        assert sorted(set(owner)) == list(range(world))
        mine = {}
        for i, (p, pt) in enumerate(zip(polys_o, points_o)):
            if owner[i] == rank:  # first half: no transcript involved (mlpcs.rs:86-97)
                pr = py.compute_pr(pt)
                s_poly = py.compute_s_polynomial(list(p), pr)
                mine[i] = (sum(a * b for a, b in zip(p, pr)) % FR, s_poly, kz.commit(s_poly))
        heads = _gather_obj({i: (v[0], v[2]) for i, v in mine.items()}, world)
        tr2, chal = py.Transcript(b"deal"), {}
        for i, (p, pt) in enumerate(zip(polys_o, points_o)):  # every rank replays the schedule of mlpcs.rs:100-105
            ev_i, sc_i = heads[owner[i]][i]
            tr2.append_fr_vec(pt)
            tr2.append_fr(ev_i)
            tr2.append_g1(sc_i)
            chal[i] = tr2.draw_field_element()
        # second halves (hyperplonk.OpeningPool.finish): the four KZG openings at r and 1/r are not absorbed by the
        # transcript, so they run after every batch's replay; S's two stay with the owner of S, the polynomial's two go
        # to the least loaded rank
        place = pool.place_poly_openings([len(p) for p in polys_o])
        assert set(place.values()) <= set(range(world)) and len(place) == 2 * len(polys_o)
        names = ("poly_opening", "poly_opening_inv", "s_opening", "s_opening_inv")
        got_o = {}
        for i, p in enumerate(polys_o):
            r, r_inv = chal[i], py.fr_inv(chal[i])
            for slot in range(4):
                who = place[(i, slot)] if slot < 2 else owner[i]
                if who == rank:
                    got_o[(i, slot)] = kz.open(p if slot < 2 else mine[i][1], r_inv if slot & 1 else r)
        merged = {}
        for d in _gather_obj(got_o, world):
            merged.update(d)
        got = [dict(evaluation_point=[x % FR for x in pt], evaluation=heads[owner[i]][i][0], s_comm=heads[owner[i]][i][1],
                    **{names[slot]: merged[(i, slot)] for slot in range(4)}) for i, pt in enumerate(points_o)]
        assert got == want_o and tr2.state == seq_tr.state
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharding_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_opening_owners_balanced_and_deterministic():
    from quill_zkvm_b200 import hyperplonk as hp
    batch = hp.OpeningBatch(None, None, None)
    sizes = [1 << 12, 1 << 12, 1 << 14, 1 << 14, 1 << 14, 1 << 14, 1 << 12, 1 << 12, 1 << 14, 1 << 14, 1 << 14]
    for n in sizes:
        batch.add(np.zeros((n, 32), np.uint8), np.zeros((3, 32), np.uint8), None)
    for w in (1, 2, 4, 8):
        owner = batch.owners(w)
        assert owner == batch.owners(w) and set(owner) <= set(range(w))
        load = [sum(s for s, o in zip(sizes, owner) if o == r) for r in range(w)]
        assert max(load) - min(load) <= max(sizes)
        # the whole proof: first halves (+ the two S openings) with their owner, the polynomials' openings pooled
        pool = hp.OpeningPool(None, None, nranks=w, rank=0)
        batch.owners(w, pool)
        place = pool.place_poly_openings(sizes)
        assert sorted(place) == [(i, s) for i in range(len(sizes)) for s in (0, 1)]
        assert max(pool.load) - min(pool.load) <= (hp.OpeningPool.BEGIN_COST + 2.0) * max(sizes)


def test_shard_ranges_partition():
    for n in (1, 7, 96, 1 << 10):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
    assert parallel.table_shard_range(5, 1, 2) == (16, 32)
    with pytest.raises(AssertionError):
        parallel.table_shard_range(1, 0, 4)


def test_sharded_view_windows():
    """hyperplonk.sharded_view: rank g's store holds the g-th contiguous 1/G of every resident table (the top variables),
    or nothing when the store does not qualify"""
    from types import SimpleNamespace

    import quill_zkvm_b200 as q
    from quill_zkvm_b200 import hyperplonk as hp

    n = 14
    store = q.VirtualPolynomialStore(n)
    store.polynomials = [q.DeviceBuffer(None, 0x1000_0000 * (t + 1), 32 << n, owner=False) for t in range(3)]
    store.virtual_polys = ["h"]
    for G in (2, 4):
        spans = []
        for rank in range(G):
            part = hp.sharded_view(SimpleNamespace(nranks=G, rank=rank), store)
            assert part is not None and part.num_vars == n and part.virtual_polys is store.virtual_polys
            for t, (p, full) in enumerate(zip(part.polynomials, store.polynomials)):
                assert not p.owner and p.nbytes == (32 << n) // G and p.ptr == full.ptr + rank * p.nbytes
            spans.append(parallel.table_shard_range(n, rank, G))
        assert spans[0][0] == 0 and spans[-1][1] == 1 << n  # the same partition the sharded entry points document
    assert hp.sharded_view(SimpleNamespace(nranks=1, rank=0), store) is None          # one rank
    assert hp.sharded_view(SimpleNamespace(nranks=8, rank=0), store) is None          # 2^11 entries per rank: too few
    host = q.VirtualPolynomialStore(n)
    host.polynomials = [np.zeros((1 << n, 32), np.uint8)]
    assert hp.sharded_view(SimpleNamespace(nranks=2, rank=0), host) is None           # host tables
