"""Callers of the hot path: logup multiset / permutation check and the HyperPlonk prove driver,
orchestrated on the host exactly as the reference does, with every heavy step on the device through the C ABI:

  MultisetEqualityProof::prove   hyperplonk/src/piops/multiset_check.rs:28-182
  PermutationCheckProof::prove   hyperplonk/src/piops/permutation_check.rs:13-58
  HyperPlonk::{preprocess,prove} hyperplonk/src/proof/proof.rs:63-301
  TransitionCircuit              hyperplonk/src/frontend/transition_circuit.rs:26-151 (builder + Circuit impl)

Host glue touches only O(#constraints) scalars (powers of the batching challenge, 0/1 column bits) with Python ints;
tables, commitments, sumchecks and openings are computed by the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .api import (Context, DeviceBuffer, EvaluationClaim, KZG, MLEvalProof, SumcheckProof, Transcript, VirtualPolyExpr,
                  VirtualPolynomialStore, ZeroCheckProof, fast_eq_eval_hypercube, logup_denominators)

FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_R = 1 << 256
_RINV = pow(_R, -1, FR)


def _host_ptr(a: np.ndarray):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a.ctypes.data


def fr_mont(v: int) -> np.ndarray:
    """int -> 32-byte Montgomery Fr (host glue for constants)."""
    return np.frombuffer(((v % FR) * _R % FR).to_bytes(32, "little"), dtype=np.uint8).copy()


def fr_int(b: np.ndarray) -> int:
    return int.from_bytes(np.ascontiguousarray(b, dtype=np.uint8).tobytes(), "little") * _RINV % FR


def fr_table(vals) -> np.ndarray:
    return np.stack([fr_mont(v) for v in vals]) if len(vals) else np.zeros((0, 32), dtype=np.uint8)


def small_int_table(ctx: Context, vals) -> np.ndarray:
    """Montgomery table of non-negative integers < 2^64 (ids, permutation indices, selector bits): canonical limbs are
    laid out with numpy and converted on the device (element-wise to_mont), avoiding a Python big-int loop."""
    v = np.ascontiguousarray(vals, dtype=np.uint64)
    raw = np.zeros((v.shape[0], 4), dtype=np.uint64)
    raw[:, 0] = v
    if v.shape[0] == 0:
        return np.zeros((0, 32), dtype=np.uint8)
    return ctx.field_op(0, 4, raw.view(np.uint8).reshape(-1, 32))


def Const(v: int) -> VirtualPolyExpr:
    return VirtualPolyExpr.Const(fr_mont(v))


def Sub(a: VirtualPolyExpr, b: VirtualPolyExpr) -> VirtualPolyExpr:
    """`a - b` on VirtualPolyExpr: Add(a, Mul(Const(-1), b)) (virtual_polynomial.rs:67-77)."""
    return a + (Const(FR - 1) * b)


def sharded_view(ctx: Context, store: VirtualPolynomialStore) -> Optional[VirtualPolynomialStore]:
    """Several ranks, every table resident on every rank: a store over this rank's contiguous 1/G of every table (the
    top log2 G variables, so every (2p, 2p+1) pair stays local: sumcheck.rs:56-57) as non-owning windows, for the
    sharded sumcheck / zero-check entry points -- instead of every rank proving the whole table.  None when the store
    does not qualify (one rank, host tables, fewer than 2^12 entries per rank)."""
    G, rank = getattr(ctx, "nranks", 1), getattr(ctx, "rank", 0)
    n = 1 << store.num_vars
    if G <= 1 or n % G or n // G < (1 << 12):
        return None
    if not store.polynomials or not all(isinstance(p, DeviceBuffer) and p.nbytes == 32 * n for p in store.polynomials):
        return None
    part = VirtualPolynomialStore(store.num_vars)
    shard = 32 * (n // G)
    part.polynomials = [p.view(rank * shard, shard) for p in store.polynomials]
    part.virtual_polys = store.virtual_polys
    return part


# ---------------------------------------------------------------------------------------------------------------------
class OpeningPool:
    """Several ranks: the second halves of ALL the MLEvalProofs of one proof, run at the end as one balanced pool.

    An MLEvalProof's four KZG openings (mlpcs.rs:109-113) are never absorbed by the transcript -- only its evaluation
    and S commitment are (mlpcs.rs:100-102) -- so nothing that follows depends on them.  A batch therefore only runs
    the first halves (dealt to the ranks), exchanges (evaluation, S commitment), replays the transcript and leaves
    (opening, r) here, where the 4 x (number of openings) KZG openings of the whole proof form one pool: the two openings
    of S stay with the rank that holds S, the two of the polynomial (resident on every rank) go to whichever rank is
    least loaded; they run when a rank would otherwise wait for a later batch's slowest first half, and the rest in
    `finish`.  With 13 openings per trace and 8 ranks, per-batch dealing left most ranks idle
    half of the time (two openings on some ranks, one on the others, traces of different sizes back to back)."""

    def __init__(self, ctx: Context, pcs: KZG, nranks: Optional[int] = None, rank: Optional[int] = None):
        self.ctx, self.pcs = ctx, pcs
        self.nranks = getattr(ctx, "nranks", 1) if nranks is None else nranks
        self.rank = getattr(ctx, "rank", 0) if rank is None else rank
        self.load = [0.0] * self.nranks  # cost dealt so far, in polynomial entries x MSMs
        self.entries, self.after = [], []
        self.place = {}    # (opening, slot) -> rank that runs that KZG opening
        self.done = {}     # (opening, slot) -> x ‖ y ‖ proof, for the openings this rank has run
        self.placed = 0    # openings whose four KZG openings have a rank

    # first half of an opening in units of one KZG opening of its length: commit(S), three transforms of twice the length,
    # the eq table and the inner product (fitted on the per-call profile of an 8-GPU proof, tools/profile_hp_multi.py)
    BEGIN_COST = 2.0

    def deal_begin(self, length: int) -> int:
        """owner of an opening's first half (and with it of its two S openings): the least loaded rank"""
        r = min(range(self.nranks), key=lambda k: (self.load[k], k))
        self.load[r] += (self.BEGIN_COST + 2.0) * max(length, 1)
        return r

    def add(self, poly, point, evaluation, s_comm, r, owner: int, pending, sink) -> None:
        self.entries.append(dict(poly=poly, point=point, ev=evaluation, sc=s_comm, r=r, owner=owner, pending=pending, sink=sink))

    def on_done(self, fn) -> None:
        self.after.append(fn)

    def place_poly_openings(self, lengths: List[int], first: int = 0) -> dict:
        """(opening, slot) -> rank for the two openings of the polynomials first, first + 1, ..: longest first onto the
        least loaded rank (pure bookkeeping: every rank takes the same decisions)"""
        place = {}
        for i in sorted(range(len(lengths)), key=lambda j: (-lengths[j], j)):
            for slot in (0, 1):
                k = min(range(self.nranks), key=lambda q: (self.load[q], q))
                self.load[k] += max(lengths[i], 1)
                place[(first + i, slot)] = k
        return place

    def place_new(self) -> None:
        """give the KZG openings of the entries added since the last call their ranks (called when a batch has replayed
        its transcript: from then on the openings can run whenever their rank has nothing better to do)"""
        new = self.entries[self.placed:]
        self.place.update(self.place_poly_openings([e["poly"].nbytes // 32 for e in new], self.placed))
        for j, e in enumerate(new):
            self.place[(self.placed + j, 2)] = self.place[(self.placed + j, 3)] = e["owner"]
        self.placed = len(self.entries)

    def run_units(self, budget: Optional[float] = None) -> None:
        """run this rank's pending KZG openings, oldest first, until `budget` (in polynomial entries; None = all) is spent"""
        ctx, pcs, spent = self.ctx, self.pcs, 0.0
        for i in range(self.placed):
            e = self.entries[i]
            for slot in range(4):
                if self.place[(i, slot)] != self.rank or (i, slot) in self.done:
                    continue
                if budget is not None and spent >= budget:
                    return
                if "xs" not in e:
                    r = np.ascontiguousarray(e["r"], dtype=np.uint8).reshape(32)
                    e["xs"] = (r, ctx.field_op(0, 3, r.reshape(1, 32)).reshape(32))  # r, 1 / r (mlpcs.rs:107)
                if slot < 2:
                    target = e["poly"]
                else:
                    pend = e["pending"]
                    target = (DeviceBuffer(ctx, pend.s_dev, 32 * pend.s_len, owner=False) if pend.s_len
                              else np.zeros((0, 32), dtype=np.uint8))
                o = pcs.open(target, e["xs"][slot & 1])
                self.done[(i, slot)] = np.concatenate([o.x, o.y, o.proof])
                spent += max(e["poly"].nbytes // 32, 1)
            pend = e["pending"]
            if pend is not None and pend.s_dev and (i, 2) in self.done and (i, 3) in self.done:
                ctx.lib.qz_dev_free(ctx.h, pend.s_dev)  # both openings of S are out: release it
                pend.s_dev = None

    def finish(self) -> None:
        ctx = self.ctx
        self.place_new()
        self.run_units()
        B = len(self.entries)
        tails = np.zeros((B, 4, 128), dtype=np.uint8)
        for (i, slot), v in self.done.items():
            tails[i, slot] = v
        all_tails = ctx.allgather(tails).reshape(self.nranks, B, 4, 128)
        for i, e in enumerate(self.entries):
            ops = np.stack([all_tails[self.place[(i, slot)], i, slot] for slot in range(4)])
            e["sink"](MLEvalProof.from_parts(e["point"].copy(), e["ev"].copy(), e["sc"].copy(), ops))
        for fn in self.after:
            fn()
        self.entries, self.after, self.place, self.done, self.placed = [], [], {}, {}, 0


class OpeningBatch:
    """A run of consecutive `MultilinearPCS::open` calls with no other transcript traffic between them (the two logup
    openings of multiset_check.rs:167-170 followed by the num_cols + num_public + 3 openings of proof.rs:202-226).

    One GPU: each opening is the fused qz_mlpcs_open, in order.  Several ranks (last row; every rank holds the
    same polynomials and runs the same transcript): an opening's evaluation, S polynomial and S commitment do not depend
    on the transcript, only its four KZG openings need the challenge r -- so the openings are dealt to the ranks, which
    run the first halves, all-gather (evaluation, S commitment) and replay the transcript schedule of mlpcs.rs:100-105
    for every opening in order; the second halves wait in the proof's OpeningPool.  Every rank ends with the complete,
    byte-identical proof."""

    def __init__(self, ctx: Context, pcs: KZG, transcript: Transcript, pool: Optional[OpeningPool] = None):
        self.ctx, self.pcs, self.transcript, self.pool = ctx, pcs, transcript, pool
        self.items, self.after = [], []

    def add(self, poly, point: np.ndarray, sink) -> None:
        """queue open(poly, point); `sink(MLEvalProof)` receives the result when the batch (or its pool) has run"""
        self.items.append((poly, np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32), sink))

    def on_done(self, fn) -> None:
        self.after.append(fn)

    @staticmethod
    def _length(poly) -> int:
        return poly.nbytes // 32

    def owners(self, nranks: int, pool: Optional[OpeningPool] = None) -> List[int]:
        """rank that runs each opening's first half: longest first onto the least loaded rank of the whole proof so far
        (`pool` carries the load of the earlier batches); ties by index / rank"""
        pool = pool if pool is not None else OpeningPool(None, None, nranks=nranks, rank=0)
        owner = [0] * len(self.items)
        for i in sorted(range(len(self.items)), key=lambda j: (-self._length(self.items[j][0]), j)):
            owner[i] = pool.deal_begin(self._length(self.items[i][0]))
        return owner

    def run(self) -> None:
        ctx, pcs, tr = self.ctx, self.pcs, self.transcript
        nranks, rank = getattr(ctx, "nranks", 1), getattr(ctx, "rank", 0)
        if nranks == 1:
            for poly, point, sink in self.items:
                sink(pcs.open_multilinear(poly, point, tr))
            for fn in self.after:
                fn()
        else:
            pool = self.pool if self.pool is not None else OpeningPool(ctx, pcs)
            B, owner = len(self.items), self.owners(nranks, pool)
            pending, head = {}, np.zeros((B, 96), dtype=np.uint8)
            for i, (poly, point, _) in enumerate(self.items):
                if owner[i] == rank:
                    pending[i] = pcs.open_multilinear_begin(poly, point)
                    head[i, :32], head[i, 32:] = pending[i].evaluation, pending[i].s_comm
            # 13 first halves over 8 ranks is two on some and one on the others, and all wait for the slowest before the
            # transcript can go on: a rank with fewer first halves spends the difference on KZG openings of EARLIER
            # batches, whose challenges are known (they would otherwise run at the end of the proof)
            begun = [0.0] * nranks
            for i, (poly, _, _) in enumerate(self.items):
                begun[owner[i]] += OpeningPool.BEGIN_COST * max(self._length(poly), 1)
            if max(begun) > begun[rank]:
                pool.run_units(max(begun) - begun[rank])
            heads = ctx.allgather(head).reshape(nranks, B, 96)
            for i, (poly, point, sink) in enumerate(self.items):  # mlpcs.rs:100-105, every rank, in order
                ev, sc = heads[owner[i], i, :32], heads[owner[i], i, 32:]
                tr.append_fr_vec(point)
                tr.append_fr(ev)
                tr.append_g1(sc)
                r = tr.draw_field_element()
                pool.add(poly, point, ev.copy(), sc.copy(), r.copy(), owner[i], pending.get(i), sink)
            pool.place_new()
            for fn in self.after:
                pool.on_done(fn)
            if self.pool is None:
                pool.finish()
        self.items, self.after = [], []


# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class MultisetEqualityProof:
    """multiset_check.rs:17-24"""
    denom_left_commitment: np.ndarray
    denom_right_commitment: np.ndarray
    sumcheck_proof: SumcheckProof
    opening_proof_denom_left: MLEvalProof
    opening_proof_denom_right: MLEvalProof

    @staticmethod
    def prove(ctx: Context, store: VirtualPolynomialStore, h_left: int, h_right: int, transcript: Transcript, pcs: KZG,
              multiplicities: Optional[int] = None, batch: Optional[OpeningBatch] = None
              ) -> Tuple["MultisetEqualityProof", np.ndarray]:
        """multiset_check.rs:28-182.  multiplicities=None is LookupMode::Equality, else Subset.  With `batch` the two
        openings are queued on it (the caller runs the batch; nothing else may touch the transcript before that)."""
        num_vars = store.num_vars
        on_dev = bool(store.polynomials) and isinstance(store.polynomials[0], DeviceBuffer)  # keep new tables where the store lives
        gamma = transcript.draw_field_element()  # :40
        left = logup_denominators(ctx, store, h_left, gamma, device=on_dev)  # :43-53
        right = logup_denominators(ctx, store, h_right, gamma, multiplicities, device=on_dev)  # :55-95
        split = getattr(ctx, "nranks", 1) > 1 and on_dev  # every rank holds the tables: each multiplies its index range
        c_left = pcs.commit_split(left) if split else pcs.commit(left)  # :98-99
        c_right = pcs.commit_split(right) if split else pcs.commit(right)
        transcript.append_g1(c_left)  # :100-101
        transcript.append_g1(c_right)
        lam = transcript.draw_field_element()  # :104-105
        alpha = transcript.draw_field_element()
        dl = store.allocate_polynomial(left)  # :108-109
        dr = store.allocate_polynomial(right)
        m = store.virtual_polys[multiplicities] if multiplicities is not None else Const(1)  # :128-131
        g = VirtualPolyExpr.Const(gamma)
        # :132-141   dl * (gamma + h_left) - 1 + lambda * (dr * (gamma + h_right) - m)
        zc = Sub(VirtualPolyExpr.Input(dl) * (g + store.virtual_polys[h_left]), Const(1)) + (
            VirtualPolyExpr.Const(lam) * Sub(VirtualPolyExpr.Input(dr) * (g + store.virtual_polys[h_right]), m))
        z = np.stack([transcript.draw_field_element() for _ in range(num_vars)]) if num_vars else np.zeros((0, 32), np.uint8)  # :144-146
        eq_tab = fast_eq_eval_hypercube(ctx, num_vars, z, device=on_dev)
        eq = store.allocate_polynomial(eq_tab)  # :149-150
        h_hat = store.new_virtual_from_expr(zc)  # :152-153
        store.mul_in_place(h_hat, eq)
        store.mul_const_in_place(h_hat, alpha)  # :156-158
        store.add_in_place(h_hat, dl)
        store.sub_in_place(h_hat, dr, fr_mont(FR - 1))
        part = sharded_view(ctx, store)
        if part is not None:  # :162-163, each rank on its shard of the tables (same bytes: tools/multi_gpu_check.py)
            sc, claim = SumcheckProof.prove(ctx, num_vars, part, h_hat, fr_mont(0), transcript, sharded=True)
        else:
            sc, claim = SumcheckProof.prove(ctx, num_vars, store, h_hat, fr_mont(0), transcript)
        point = claim.point
        proof = MultisetEqualityProof(c_left, c_right, sc, None, None)
        own_batch = batch is None
        if own_batch:
            batch = OpeningBatch(ctx, pcs, transcript)
        batch.add(left, point, lambda o: setattr(proof, "opening_proof_denom_left", o))  # :167-170
        batch.add(right, point, lambda o: setattr(proof, "opening_proof_denom_right", o))
        if on_dev:
            batch.on_done(lambda: [b.free() for b in (left, right, eq_tab)])
        if own_batch:
            batch.run()
        return proof, point


@dataclass
class PermutationCheckProof:
    """permutation_check.rs:8-10"""
    multiset_equality_proof: MultisetEqualityProof

    @staticmethod
    def prove(ctx: Context, store: VirtualPolynomialStore, h_left: int, h_right: int, id_indices: np.ndarray,
              permutation_indices: np.ndarray, transcript: Transcript, pcs: KZG, batch: Optional[OpeningBatch] = None):
        """permutation_check.rs:13-58"""
        assert id_indices.nbytes == 32 << store.num_vars and permutation_indices.nbytes == 32 << store.num_vars
        id_ref = store.allocate_polynomial(id_indices)  # :27-28
        perm_ref = store.allocate_polynomial(permutation_indices)
        alpha = transcript.draw_field_element()  # :30
        hl = store.new_virtual_from_virtual(h_left)  # :33-35  id + alpha * h_left
        store.mul_const_in_place(hl, alpha)
        store.add_in_place(hl, id_ref)
        hr = store.new_virtual_from_virtual(h_right)  # :38-40
        store.mul_const_in_place(hr, alpha)
        store.add_in_place(hr, perm_ref)
        proof, point = MultisetEqualityProof.prove(ctx, store, hl, hr, transcript, pcs, batch=batch)  # :42-50
        return PermutationCheckProof(proof), point


# ---------------------------------------------------------------------------------------------------------------------
class TransitionCircuit:
    """frontend/transition_circuit.rs:26-151: state cells copied row i -> i+1, recurring and boundary constraints."""

    def __init__(self, num_rows: int):
        self.num_columns = 0
        self.num_rows_ = num_rows
        self.state_cells: List[Tuple[int, int]] = []
        self.recurring_constraints: List[VirtualPolyExpr] = []
        self.boundary_constraints: List[Tuple[int, VirtualPolyExpr]] = []

    def allocate_witness_cell(self) -> int:
        self.num_columns += 1
        return self.num_columns - 1

    def allocate_state_cell(self) -> Tuple[int, int]:
        cur, nxt = self.allocate_witness_cell(), self.allocate_witness_cell()
        self.state_cells.append((cur, nxt))
        return cur, nxt

    def enforce_constraint(self, e: VirtualPolyExpr):
        self.recurring_constraints.append(e)

    def enforce_boundary_constraint(self, row: int, e: VirtualPolyExpr):
        self.boundary_constraints.append((row, e))

    # Circuit trait (proof/circuit.rs:21-74)
    def num_rows(self) -> int:
        return self.num_rows_

    def num_cols(self) -> int:
        return 1 << max(self.num_columns - 1, 0).bit_length() if self.num_columns > 1 else 1  # next_power_of_two

    def num_public_columns(self) -> int:
        return len(self.boundary_constraints)

    def public_values(self) -> List[np.ndarray]:  # :96-102 (one 0/1 selector column per boundary constraint)
        pub = [np.zeros(self.num_rows_, dtype=np.uint64) for _ in self.boundary_constraints]
        for i, (row, _) in enumerate(self.boundary_constraints):
            pub[i][row] = 1
        return pub

    def zero_check_expressions(self) -> List[VirtualPolyExpr]:  # :104-122
        out = list(self.recurring_constraints)
        for i, (_row, c) in enumerate(self.boundary_constraints):
            out.append(VirtualPolyExpr.Input(i + self.num_cols()) * c)
        return out

    def permutation(self) -> Tuple[np.ndarray, np.ndarray]:  # :124-151
        n, rows = self.num_rows_ * self.num_cols(), self.num_rows_
        assert n & (n - 1) == 0
        perm = np.arange(n, dtype=np.uint64)
        r = np.arange(rows - 1, dtype=np.uint64)
        for cur, nxt in self.state_cells:  # (next, row) <-> (current, row + 1), applied in order like the reference's loop
            frm, to = nxt * rows + r, cur * rows + r + 1
            perm[frm] = to
            perm[to] = frm
        return np.arange(1, n + 1, dtype=np.uint64), perm + 1  # +1: no zero values


@dataclass
class TraceProof:
    """proof/proof.rs:17-25"""
    zero_check_proof: ZeroCheckProof
    permutation_check_proof: PermutationCheckProof
    openings_zero_check: List[MLEvalProof]
    openings_public: List[MLEvalProof]
    opening_id: MLEvalProof
    opening_permutation: MLEvalProof
    opening_permutation_trace: MLEvalProof


@dataclass
class HyperPlonkProof:
    witness_commitment: List[np.ndarray]
    trace_proofs: List[TraceProof]
    transcript_state: bytes = b""  # final Fiat-Shamir state (not in the reference struct; exposed for parity checks)


@dataclass
class _TracePK:
    id_poly: DeviceBuffer
    permutation_poly: DeviceBuffer
    public_values: List[np.ndarray]
    public_columns: List[DeviceBuffer] = None  # the un-padded public columns, resident (circuit constants: built once)


@dataclass
class _TraceVK:
    circuit: TransitionCircuit
    public_columns_commitments: List[np.ndarray]
    id_commitment: np.ndarray
    permutation_commitment: np.ndarray


class HyperPlonk:
    """proof/proof.rs:12-15, 63-301"""

    def __init__(self, ctx: Context, trace_pks, trace_vks):
        self.ctx, self.trace_pks, self.trace_vks = ctx, trace_pks, trace_vks

    @staticmethod
    def preprocess(ctx: Context, circuits: List[TransitionCircuit], pcs: KZG) -> "HyperPlonk":
        pks, vks = [], []
        for c in circuits:  # preprocess_trace, :63-122
            rows, cols = c.num_rows(), c.num_cols()
            assert rows & (rows - 1) == 0, "Number of rows must be a power of two"
            assert cols & (cols - 1) == 0, "Number of columns must be a power of two"
            n = rows * cols
            pub = []
            for col in c.public_values():
                assert len(col) == rows, "Public column length mismatch"
                pub.append(small_int_table(ctx, np.concatenate([col, np.zeros(n - rows, dtype=np.uint64)])))  # :77-86
            pub_comms = [pcs.commit(p) for p in pub]
            ids, perm = c.permutation()
            assert len(ids) == n and len(perm) == n
            id_t, perm_t = ctx.upload(small_int_table(ctx, ids)), ctx.upload(small_int_table(ctx, perm))  # resident for prove()
            vks.append(_TraceVK(c, pub_comms, pcs.commit(id_t), pcs.commit(perm_t)))
            pks.append(_TracePK(id_t, perm_t, pub, [ctx.upload(small_int_table(ctx, col)) for col in c.public_values()]))
        return HyperPlonk(ctx, pks, vks)

    def _prove_trace(self, pcs: KZG, full_witness: DeviceBuffer, transcript: Transcript, pk: _TracePK,
                     circuit: TransitionCircuit, pool: Optional[OpeningPool] = None) -> TraceProof:
        """prove_trace (:145-237).  The trace stays in HBM: `full_witness` is the column-major witness on the device and
        every column handed to the zero-check is a window of it."""
        ctx = self.ctx
        log2_rows, log2_cols = circuit.num_rows().bit_length() - 1, circuit.num_cols().bit_length() - 1
        col_bytes = 32 * circuit.num_rows()
        store = VirtualPolynomialStore(log2_rows)  # :156-162
        for col in range(circuit.num_cols()):
            store.allocate_polynomial(full_witness.view(col * col_bytes, col_bytes))
        public = pk.public_columns  # proof.rs:159-162 rebuilds them per proof; they are constants of the circuit
        for p in public:
            store.allocate_polynomial(p)
        exprs = circuit.zero_check_expressions()  # :165-175
        alpha = fr_int(transcript.draw_field_element())
        zc_expr = Const(0)
        for i, e in enumerate(exprs):
            zc_expr = zc_expr + (Const(pow(alpha, i, FR)) * e)
        zc_virtual = store.new_virtual_from_expr(zc_expr)
        part = sharded_view(ctx, store)
        if part is not None:  # :178-180, each rank on its shard of the columns
            zero_check_proof, zc_claim = ZeroCheckProof.prove(ctx, part, zc_virtual, transcript, sharded=True)
        else:
            zero_check_proof, zc_claim = ZeroCheckProof.prove(ctx, store, zc_virtual, transcript)
        store2 = VirtualPolynomialStore(log2_rows + log2_cols)  # :184-196
        w_idx = store2.allocate_polynomial(full_witness)
        w_virtual = store2.new_virtual_from_input(w_idx)
        # every opening of the trace forms one run in the transcript: queue them, then run the batch (split over the
        # ranks when there are several)
        batch = OpeningBatch(ctx, pcs, transcript, pool)
        perm_proof, perm_point = PermutationCheckProof.prove(ctx, store2, w_virtual, w_virtual, pk.id_poly,
                                                             pk.permutation_poly, transcript, pcs, batch=batch)
        openings_zc = [None] * circuit.num_cols()  # :202-210: column bits appended as the HIGH variables of the witness
        for col in range(circuit.num_cols()):
            bits = fr_table([(col >> i) & 1 for i in range(log2_cols)])
            point = np.concatenate([zc_claim.point, bits.reshape(-1, 32)]) if log2_cols else zc_claim.point
            batch.add(full_witness, point, lambda o, col=col: openings_zc.__setitem__(col, o))
        openings_pub = [None] * circuit.num_public_columns()  # :214-219 (un-padded columns)
        for i in range(circuit.num_public_columns()):
            batch.add(public[i], zc_claim.point, lambda o, i=i: openings_pub.__setitem__(i, o))
        tail = {}
        batch.add(pk.id_poly, perm_point, lambda o: tail.__setitem__("id", o))  # :222-226
        batch.add(pk.permutation_poly, perm_point, lambda o: tail.__setitem__("perm", o))
        batch.add(full_witness, perm_point, lambda o: tail.__setitem__("trace", o))
        batch.run()
        proof = TraceProof(zero_check_proof, perm_proof, openings_zc, openings_pub, tail.get("id"), tail.get("perm"), tail.get("trace"))
        if pool is not None and getattr(ctx, "nranks", 1) > 1:  # the openings arrive when the pool has run
            def fill():
                proof.opening_id, proof.opening_permutation, proof.opening_permutation_trace = tail["id"], tail["perm"], tail["trace"]
            pool.on_done(fill)
        return proof

    def prove(self, pcs: KZG, witness_traces: List[List[np.ndarray]]) -> HyperPlonkProof:
        """proof.rs:239-301.  witness_traces[t][c] is column c of trace t as a (rows, 32) Montgomery array."""
        ctx = self.ctx
        transcript = Transcript(b"hyperplonk_proof", ctx)  # :245
        multi = getattr(ctx, "nranks", 1) > 1
        pool = OpeningPool(ctx, pcs) if multi else None
        comms, fulls = [], []
        for witness, vk in zip(witness_traces, self.trace_vks):  # :250-284
            c = vk.circuit
            assert len(witness) == c.num_cols(), "Witness columns length mismatch"
            for col in witness:
                assert col.shape[0] == c.num_rows(), "Witness column row length mismatch"
            full = ctx.alloc(32 * c.num_cols() * c.num_rows())  # column-major (:270), uploaded once and kept resident
            for j, col in enumerate(witness):
                ctx.check(ctx.lib.qz_dev_upload(ctx.h, full.ptr + j * 32 * c.num_rows(), _host_ptr(col), 32 * c.num_rows()))
            com = pcs.commit_split(full) if multi else pcs.commit(full)
            transcript.append_g1(com)
            comms.append(com)
            fulls.append(full)
        proofs = [self._prove_trace(pcs, fulls[i], transcript, self.trace_pks[i], self.trace_vks[i].circuit, pool)
                  for i in range(len(witness_traces))]
        if pool is not None:
            pool.finish()
        for f in fulls:
            f.free()
        return HyperPlonkProof(comms, proofs, transcript.state.tobytes())
