"""Host-side mirror of the reference's Rust API for the hot path, over the C ABI.

Names, argument meaning and error behaviour follow the reference so the parity tests read like its own tests:
  Transcript                 transcript/src/transcript.rs:14-75
  KZG.commit / KZG.open      pcs/src/kzg.rs:61-96
  VirtualPolyExpr / Store    hyperplonk/src/utils/virtual_polynomial.rs:9-18, 142-278
  SumcheckProof.prove        hyperplonk/src/piops/sumcheck.rs:28-114
  ZeroCheckProof.prove       hyperplonk/src/piops/zerocheck.rs:14-49
  fast_eq_eval_hypercube     hyperplonk/src/utils/eq_eval.rs:6-31

Field elements are numpy uint8 arrays of 32 little-endian Montgomery bytes (the memory layout of ark_bn254::Fr);
G1 points are 64 bytes x‖y (Montgomery), all-zero = infinity.  Python is plumbing only: nothing here does field or
curve arithmetic.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import QuillError


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _u8(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a.reshape(shape) if shape is not None else a


class Context:
    """One device + stream + scratch (qz_ctx).  `stream` may be a raw cudaStream_t (e.g. torch's current stream)."""

    def __init__(self, device: int = 0, stream: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.qz_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc:
            raise QuillError(rc, self.lib.qz_status_str(rc).decode())
        self.h = h
        self.device = device

    def check(self, rc: int):
        if rc:
            raise QuillError(rc, f"{self.lib.qz_status_str(rc).decode()}: {self.lib.qz_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.qz_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(self.lib.qz_ctx_sync(self.h))

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.qz_kernel_launches(self.h))

    def last_elapsed_ms(self, which: int = 0) -> float:
        return float(self.lib.qz_last_elapsed_ms(self.h, which))

    def last_stat(self, which: int) -> float:
        return float(self.lib.qz_last_stat(self.h, which))

    def msm_accumulate_stats(self, reset: int = 0):
        """(summed msm_accumulate kernel ms, mixed additions, launches) since collection was started with reset=1;
        reset=-1 reads and stops (qz_msm_accumulate_stats)."""
        ms, adds, n = C.c_double(), C.c_double(), C.c_uint64()
        self.check(self.lib.qz_msm_accumulate_stats(self.h, reset, C.byref(ms), C.byref(adds), C.byref(n)))
        return ms.value, adds.value, int(n.value)

    # -- device buffers ------------------------------------------------------------------------------------------
    def alloc(self, nbytes: int) -> "DeviceBuffer":
        p = C.c_void_p()
        self.check(self.lib.qz_dev_alloc(self.h, nbytes, C.byref(p)))
        return DeviceBuffer(self, p.value, nbytes)

    def trim(self):
        """Return the device blocks parked by DeviceBuffer.free() to the driver (they are otherwise reused)."""
        self.check(self.lib.qz_dev_trim(self.h))

    def upload(self, host: np.ndarray) -> "DeviceBuffer":
        host = _u8(host)
        buf = self.alloc(host.nbytes)
        self.check(self.lib.qz_dev_upload(self.h, buf.ptr, _ptr(host), host.nbytes))
        return buf

    def random_fr(self, n: int, seed: int) -> "DeviceBuffer":
        buf = self.alloc(32 * n)
        self.check(self.lib.qz_dev_random_fr(self.h, buf.ptr, n, seed))
        return buf

    # -- measurement ------------------------------------------------------------------------------------------------
    def bench_imad(self, variant: int = 0) -> float:
        out = C.c_double()
        self.check(self.lib.qz_bench_imad(self.h, variant, C.byref(out)))
        return out.value

    def bench_fp_mul(self, field: int = 1) -> float:
        out = C.c_double()
        self.check(self.lib.qz_bench_fp_mul(self.h, field, C.byref(out)))
        return out.value

    # -- multi-GPU ----------------------------------------------------------------------------------------------------
    def comm_unique_id(self) -> np.ndarray:
        out = np.zeros(128, dtype=np.uint8)
        rc = self.lib.qz_comm_unique_id(_ptr(out))
        if rc:
            raise QuillError(rc, "ncclGetUniqueId")
        return out

    def comm_init(self, unique_id: np.ndarray, rank: int, nranks: int):
        self.check(self.lib.qz_comm_init(self.h, _ptr(_u8(unique_id)), rank, nranks))
        self.rank, self.nranks = rank, nranks

    @property
    def peer_memory(self) -> bool:
        """True when the ranks exchange through mailboxes in each other's HBM (CUDA IPC) instead of NCCL all-gathers."""
        return bool(self.lib.qz_comm_peer_memory(self.h))

    def comm_resync(self) -> None:
        """Collective: re-agree on the mailbox sequence numbers after a sharded call failed on some rank (qz_comm_resync)."""
        self.check(self.lib.qz_comm_resync(self.h))

    def allgather(self, mine: np.ndarray) -> np.ndarray:
        """All-gather a small uint8 array over the library's communicator -> (nranks, len)."""
        mine = _u8(mine).reshape(-1)
        n = getattr(self, "nranks", 1)
        out = np.zeros((n, mine.shape[0]), dtype=np.uint8)
        if n == 1:
            out[0] = mine
        elif mine.shape[0]:
            self.check(self.lib.qz_comm_allgather_host(self.h, _ptr(mine), _ptr(out), mine.shape[0]))
        return out

    # -- test hooks ----------------------------------------------------------------------------------------------------
    def field_op(self, field: int, op: int, a: np.ndarray, b: Optional[np.ndarray] = None) -> np.ndarray:
        a = _u8(a, (-1, 32))
        out = np.zeros_like(a)
        bp = _ptr(_u8(b, (-1, 32))) if b is not None else None
        self.check(self.lib.qz_test_field_op(self.h, field, op, _ptr(a), bp, _ptr(out), a.shape[0]))
        return out

    def fold(self, r: np.ndarray, a0: np.ndarray, a1: np.ndarray) -> np.ndarray:
        """test hook: a0 + r (a1 - a0) element-wise through the fixed-challenge fold of the large sumcheck passes"""
        a0, a1 = _u8(a0).reshape(-1, 32), _u8(a1).reshape(-1, 32)
        out = np.zeros_like(a0)
        self.check(self.lib.qz_test_fold(self.h, _ptr(_u8(r, (32,))), _ptr(a0), _ptr(a1), _ptr(out), a0.shape[0]))
        return out

    def g1_add(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a, b = _u8(a, (-1, 64)), _u8(b, (-1, 64))
        out = np.zeros_like(a)
        self.check(self.lib.qz_test_g1_add(self.h, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        return out

    def g1_mul(self, a: np.ndarray, scalars: np.ndarray) -> np.ndarray:
        a, s = _u8(a, (-1, 64)), _u8(scalars, (-1, 32))
        out = np.zeros_like(a)
        self.check(self.lib.qz_test_g1_mul(self.h, _ptr(a), _ptr(s), _ptr(out), a.shape[0]))
        return out

    def g1_serialize(self, xy: np.ndarray) -> bytes:
        out = np.zeros(64, dtype=np.uint8)
        self.check(self.lib.qz_g1_serialize(self.h, _ptr(_u8(xy, (64,))), _ptr(out)))
        return out.tobytes()


class DeviceBuffer:
    def __init__(self, ctx: Context, ptr: int, nbytes: int, owner: bool = True):
        self.ctx, self.ptr, self.nbytes, self.owner = ctx, ptr, nbytes, owner

    def view(self, offset: int, nbytes: int) -> "DeviceBuffer":
        """Non-owning window [offset, offset + nbytes) of this buffer (e.g. one column of a column-major witness)."""
        assert 0 <= offset and offset + nbytes <= self.nbytes
        return DeviceBuffer(self.ctx, self.ptr + offset, nbytes, owner=False)

    def download(self) -> np.ndarray:
        out = np.zeros(self.nbytes, dtype=np.uint8)
        self.ctx.check(self.ctx.lib.qz_dev_download(self.ctx.h, _ptr(out), self.ptr, self.nbytes))
        return out

    def free(self):
        if self.ptr and self.owner:
            self.ctx.lib.qz_dev_free(self.ctx.h, self.ptr)
        self.ptr = None


def _table_ptr(t):
    """(pointer, on_device) of a table given as DeviceBuffer or numpy array."""
    if isinstance(t, DeviceBuffer):
        return t.ptr, True
    return t.ctypes.data, False


# ---------------------------------------------------------------------------------------------------------------------
class Transcript:
    """transcript/src/transcript.rs:14-75.  `state` is the 32-byte blake3 state (a pub field in the reference)."""

    def __init__(self, domain: bytes, ctx: Optional[Context] = None):
        self.lib = _lib.load()
        self.ctx = ctx
        self.domain = bytes(domain)
        self.state = np.zeros(32, dtype=np.uint8)
        d = np.frombuffer(self.domain, dtype=np.uint8).copy() if self.domain else np.zeros(1, dtype=np.uint8)
        self.lib.qz_transcript_new(_ptr(d), len(self.domain), _ptr(self.state))

    def append_bytes(self, message: bytes):
        m = np.frombuffer(bytes(message), dtype=np.uint8).copy() if message else np.zeros(1, dtype=np.uint8)
        self.lib.qz_transcript_append_bytes(_ptr(self.state), _ptr(m), len(message))

    def append_usize(self, v: int):
        """append_serializable(&usize): u64 little-endian."""
        self.append_bytes(int(v).to_bytes(8, "little"))

    def append_fr(self, fr: np.ndarray):
        """append_serializable(&Fr): 32-byte canonical little-endian (conversion from Montgomery on the device)."""
        self.ctx.check(self.lib.qz_transcript_append_fr(self.ctx.h, _ptr(self.state), _ptr(_u8(fr, (32,)))))

    def append_fr_vec(self, frs: np.ndarray):
        """append_serializable(&[Fr]) / &Vec<Fr>: u64 LE length, then each element canonical (mlpcs.rs:100)."""
        frs = _u8(frs, (-1, 32))
        body = self.ctx.field_op(0, 5, frs).tobytes() if frs.shape[0] else b""  # Montgomery -> canonical on the device
        self.append_bytes(int(frs.shape[0]).to_bytes(8, "little") + body)

    def append_g1(self, xy: np.ndarray):
        """append_serializable(&G1): ark-serialize uncompressed encoding."""
        self.ctx.check(self.lib.qz_transcript_append_g1(self.ctx.h, _ptr(self.state), _ptr(_u8(xy, (64,)))))

    def draw_challenge(self, n: int) -> bytes:
        out = np.zeros(n, dtype=np.uint8)
        self.lib.qz_transcript_draw_challenge(_ptr(self.state), _ptr(out), n)
        return out.tobytes()

    def draw_field_element(self) -> np.ndarray:
        out = np.zeros(32, dtype=np.uint8)
        self.ctx.check(self.lib.qz_transcript_draw_fr(self.ctx.h, _ptr(self.state), _ptr(out)))
        return out


@dataclass
class EvaluationClaim:
    """pcs/src/lib.rs:10-13."""
    point: np.ndarray  # (n, 32)
    evaluation: np.ndarray  # (32,)


# ---------------------------------------------------------------------------------------------------------------------
class SRS:
    """Device-resident affine G1 powers (qz_srs)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    def __len__(self):
        return int(self.ctx.lib.qz_srs_len(self.h))

    def download(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = len(self) - first if count is None else count
        out = np.zeros((count, 64), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.qz_srs_download(self.ctx.h, self.h, first, count, _ptr(out)))
        return out

    def free(self):
        if self.h:
            self.ctx.lib.qz_srs_free(self.h)
            self.h = None


@dataclass
class KZGOpeningProof:
    """pcs/src/kzg.rs:25-32."""
    x: np.ndarray
    y: np.ndarray
    proof: np.ndarray  # G1, 64 B


@dataclass
class MLEvalProof:
    """pcs/src/mlpcs.rs:32-44."""
    evaluation_point: np.ndarray
    evaluation: np.ndarray
    s_comm: np.ndarray
    poly_opening: KZGOpeningProof
    poly_opening_inv: KZGOpeningProof
    s_opening: KZGOpeningProof
    s_opening_inv: KZGOpeningProof

    @staticmethod
    def from_parts(eval_point, evaluation, s_comm, ops) -> "MLEvalProof":
        """ops: (4, 128) rows of x ‖ y ‖ proof in the order of mlpcs.rs:109-113"""
        o = [KZGOpeningProof(ops[i, :32].copy(), ops[i, 32:64].copy(), ops[i, 64:].copy()) for i in range(4)]
        return MLEvalProof(eval_point, evaluation, s_comm, o[0], o[1], o[2], o[3])


@dataclass
class PendingOpening:
    """An MLEvalProof between its two halves (KZG.open_multilinear_begin / _finish)."""
    poly: object
    eval_point: np.ndarray
    evaluation: np.ndarray
    s_comm: np.ndarray
    s_dev: Optional[int]
    s_len: int


class KZG:
    """pcs/src/kzg.rs:10-96, prover side.  `g1_points` live on the device, normalised once."""

    def __init__(self, ctx: Context, srs: SRS):
        self.ctx, self.srs = ctx, srs
        self.max_degree = len(srs) - 1

    @classmethod
    def from_points(cls, ctx: Context, xy: np.ndarray) -> "KZG":
        xy = _u8(xy, (-1, 64))
        h = C.c_void_p()
        ctx.check(ctx.lib.qz_srs_upload(ctx.h, _ptr(xy), xy.shape[0], C.byref(h)))
        return cls(ctx, SRS(ctx, h))

    @classmethod
    def trusted_setup(cls, ctx: Context, max_degree: int, g1: np.ndarray, tau: np.ndarray) -> "KZG":
        """kzg.rs:35-59 with the generator and tau supplied by the caller (the reference draws them from an RNG)."""
        h = C.c_void_p()
        ctx.check(ctx.lib.qz_srs_generate(ctx.h, _ptr(_u8(g1, (64,))), _ptr(_u8(tau, (32,))), max_degree + 1, C.byref(h)))
        return cls(ctx, SRS(ctx, h))

    def precompute(self, window_bits: int = 0) -> "KZG":
        """One-time: store the window multiples 2^(c w) P_i on the device (qz_srs_precompute)."""
        self.ctx.check(self.ctx.lib.qz_srs_precompute(self.ctx.h, self.srs.h, window_bits))
        return self

    def commit(self, polynomial) -> np.ndarray:
        """kzg.rs:61-73.  Raises AssertionError like the reference's assert! when the polynomial is too long."""
        ptr, on_dev, n = self._coeffs(polynomial)
        out = np.zeros(64, dtype=np.uint8)
        rc = self.ctx.lib.qz_kzg_commit(self.ctx.h, self.srs.h, ptr, n, int(on_dev), _ptr(out))
        if rc == _lib.QZ_ERR_DEGREE:
            raise AssertionError("Polynomial degree exceeds max degree")
        self.ctx.check(rc)
        return out

    def msm_unchecked(self, scalars) -> np.ndarray:
        """VariableBaseMSM::msm_unchecked on the SRS (zips to the shorter input)."""
        ptr, on_dev, n = self._coeffs(scalars)
        out = np.zeros(64, dtype=np.uint8)
        self.ctx.check(self.ctx.lib.qz_msm(self.ctx.h, self.srs.h, ptr, n, int(on_dev), _ptr(out)))
        return out

    def msm_sharded(self, scalars) -> np.ndarray:
        ptr, on_dev, n = self._coeffs(scalars)
        out = np.zeros(64, dtype=np.uint8)
        self.ctx.check(self.ctx.lib.qz_msm_sharded(self.ctx.h, self.srs.h, ptr, n, int(on_dev), _ptr(out)))
        return out

    def commit_split(self, coeffs) -> np.ndarray:
        """KZG::commit when every rank holds the whole SRS and the whole polynomial: each multiplies its index range and
        all receive the commitment (qz_msm_split); on one rank it is `commit`."""
        ptr, on_dev, n = self._coeffs(coeffs)
        out = np.zeros(64, dtype=np.uint8)
        rc = self.ctx.lib.qz_msm_split(self.ctx.h, self.srs.h, ptr, n, int(on_dev), _ptr(out))
        if rc == _lib.QZ_ERR_DEGREE:
            raise AssertionError("Polynomial degree exceeds max degree")
        self.ctx.check(rc)
        return out

    def open(self, polynomial, x: np.ndarray) -> KZGOpeningProof:
        """kzg.rs:75-96."""
        ptr, on_dev, n = self._coeffs(polynomial)
        y = np.zeros(32, dtype=np.uint8)
        pr = np.zeros(64, dtype=np.uint8)
        x = _u8(x, (32,))
        rc = self.ctx.lib.qz_kzg_open(self.ctx.h, self.srs.h, ptr, n, int(on_dev), _ptr(x), _ptr(y), _ptr(pr))
        if rc == _lib.QZ_ERR_DEGREE:
            raise AssertionError("Polynomial degree exceeds max degree")
        self.ctx.check(rc)
        return KZGOpeningProof(x.copy(), y, pr)

    def open_multilinear(self, poly, eval_point: np.ndarray, transcript: "Transcript") -> "MLEvalProof":
        """MultilinearPCS::open (pcs/src/mlpcs.rs:191-198) = MLEvalProof::prove (mlpcs.rs:83-124)."""
        ptr, on_dev, n = self._coeffs(poly)
        pt = _u8(eval_point, (-1, 32))
        ev = np.zeros(32, dtype=np.uint8)
        sc = np.zeros(64, dtype=np.uint8)
        ops = np.zeros((4, 128), dtype=np.uint8)
        rc = self.ctx.lib.qz_mlpcs_open(self.ctx.h, self.srs.h, ptr, n, int(on_dev), _ptr(pt) if pt.shape[0] else None,
                                        pt.shape[0], _ptr(transcript.state), _ptr(ev), _ptr(sc), _ptr(ops))
        if rc == _lib.QZ_ERR_DEGREE:
            raise AssertionError("Polynomial degree exceeds max degree")
        self.ctx.check(rc)
        return MLEvalProof.from_parts(pt.copy(), ev, sc, ops)

    def open_multilinear_begin(self, poly, eval_point: np.ndarray) -> "PendingOpening":
        """First half of MLEvalProof::prove (mlpcs.rs:86-97): evaluation, S (kept on the device), commit(S)."""
        ptr, on_dev, n = self._coeffs(poly)
        pt = _u8(eval_point, (-1, 32))
        ev = np.zeros(32, dtype=np.uint8)
        sc = np.zeros(64, dtype=np.uint8)
        s_dev, s_len = C.c_void_p(), C.c_size_t()
        rc = self.ctx.lib.qz_mlpcs_open_begin(self.ctx.h, self.srs.h, ptr, n, int(on_dev), _ptr(pt) if pt.shape[0] else None,
                                              pt.shape[0], _ptr(ev), _ptr(sc), C.byref(s_dev), C.byref(s_len))
        if rc == _lib.QZ_ERR_DEGREE:
            raise AssertionError("Polynomial degree exceeds max degree")
        self.ctx.check(rc)
        return PendingOpening(poly, pt.copy(), ev, sc, s_dev.value, s_len.value)

    def open_multilinear_finish(self, pending: "PendingOpening", r: np.ndarray) -> "MLEvalProof":
        """Second half (mlpcs.rs:107-124): the four KZG openings at r and 1/r; releases S."""
        ptr, on_dev, n = self._coeffs(pending.poly)
        ops = np.zeros((4, 128), dtype=np.uint8)
        rc = self.ctx.lib.qz_mlpcs_open_finish(self.ctx.h, self.srs.h, ptr, n, int(on_dev), pending.s_dev, pending.s_len,
                                               _ptr(_u8(r, (32,))), _ptr(ops))
        self.ctx.lib.qz_dev_free(self.ctx.h, pending.s_dev)
        pending.s_dev = None
        if rc == _lib.QZ_ERR_DEGREE:
            raise AssertionError("Polynomial degree exceeds max degree")
        self.ctx.check(rc)
        return MLEvalProof.from_parts(pending.eval_point, pending.evaluation, pending.s_comm, ops)

    def compute_s_polynomial(self, p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
        """InnerProductProof::compute_s_polynomial (pcs/src/ipa.rs:122-157), coefficients not trimmed."""
        p1, p2 = _u8(p1, (-1, 32)), _u8(p2, (-1, 32))
        L = max(p1.shape[0], p2.shape[0])
        out = np.zeros((max(L - 1, 0), 32), dtype=np.uint8)
        if L >= 2:
            self.ctx.check(self.ctx.lib.qz_compute_s_polynomial(self.ctx.h, _ptr(p1) if p1.shape[0] else None, p1.shape[0],
                                                                _ptr(p2) if p2.shape[0] else None, p2.shape[0], _ptr(out)))
        return out

    def _coeffs(self, p):
        if isinstance(p, DeviceBuffer):
            return p.ptr, True, p.nbytes // 32
        self._keep = _u8(p, (-1, 32))
        return (self._keep.ctypes.data if self._keep.shape[0] else None), False, self._keep.shape[0]


# ---------------------------------------------------------------------------------------------------------------------
class VirtualPolyExpr:
    """virtual_polynomial.rs:9-18 as an immutable tree; +, *, - build Add / Mul / Add(l, Mul(Const(-1), r))."""

    __slots__ = ("op", "a", "b", "c")

    def __init__(self, op, a=None, b=None, c=None):
        self.op, self.a, self.b, self.c = op, a, b, c

    @staticmethod
    def Input(i: int) -> "VirtualPolyExpr":
        return VirtualPolyExpr("in", i)

    @staticmethod
    def Const(c_mont: np.ndarray) -> "VirtualPolyExpr":
        return VirtualPolyExpr("const", c=_u8(c_mont, (32,)).copy())

    def __add__(self, o):
        return VirtualPolyExpr("add", self, o)

    def __mul__(self, o):
        return VirtualPolyExpr("mul", self, o)

    def sub(self, o, minus_one: np.ndarray):
        """virtual_polynomial.rs:67-77: self + Const(-1) * o (the caller supplies -1 in Montgomery form)."""
        return VirtualPolyExpr("add", self, VirtualPolyExpr("mul", VirtualPolyExpr.Const(minus_one), o))

    def flatten(self):
        """-> (nodes (m,3) uint32 children-before-parents root-last, consts (c,32) uint8)."""
        nodes: List[tuple] = []
        consts: List[np.ndarray] = []

        def go(e) -> int:
            if e.op == "in":
                nodes.append((0, e.a, 0))
            elif e.op == "const":
                consts.append(e.c)
                nodes.append((1, len(consts) - 1, 0))
            else:
                l = go(e.a)
                r = go(e.b)
                nodes.append((2 if e.op == "add" else 3, l, r))
            return len(nodes) - 1

        go(self)
        nd = np.array(nodes, dtype=np.uint32).reshape(-1, 3)
        cs = np.stack(consts).astype(np.uint8) if consts else np.zeros((0, 32), dtype=np.uint8)
        return np.ascontiguousarray(nd), np.ascontiguousarray(cs)


class VirtualPolynomialStore:
    """virtual_polynomial.rs:142-278: multilinear tables + expression trees over them."""

    def __init__(self, num_vars: int):
        self.num_vars = num_vars
        self.polynomials: list = []  # numpy (2^n, 32) arrays or DeviceBuffers
        self.virtual_polys: List[VirtualPolyExpr] = []

    def allocate_polynomial(self, poly_evals) -> int:
        n = poly_evals.nbytes // 32
        assert n == 1 << self.num_vars, "Input polynomial evaluations length does not match number of variables"
        if not isinstance(poly_evals, DeviceBuffer):
            poly_evals = _u8(poly_evals, (-1, 32))
        self.polynomials.append(poly_evals)
        return len(self.polynomials) - 1

    def new_virtual_from_input(self, g: int) -> int:
        self.virtual_polys.append(VirtualPolyExpr.Input(g))
        return len(self.virtual_polys) - 1

    def new_virtual_from_expr(self, e: VirtualPolyExpr) -> int:
        self.virtual_polys.append(e)
        return len(self.virtual_polys) - 1

    def new_virtual_from_virtual(self, v: int) -> int:
        self.virtual_polys.append(self.virtual_polys[v])
        return len(self.virtual_polys) - 1

    def add_in_place(self, f: int, g: int):
        self.virtual_polys[f] = self.virtual_polys[f] + VirtualPolyExpr.Input(g)

    def add_const_in_place(self, f: int, c: np.ndarray):
        self.virtual_polys[f] = self.virtual_polys[f] + VirtualPolyExpr.Const(c)

    def mul_in_place(self, f: int, g: int):
        self.virtual_polys[f] = self.virtual_polys[f] * VirtualPolyExpr.Input(g)

    def mul_const_in_place(self, f: int, c: np.ndarray):
        self.virtual_polys[f] = self.virtual_polys[f] * VirtualPolyExpr.Const(c)

    def sub_in_place(self, f: int, g: int, minus_one: np.ndarray):
        self.virtual_polys[f] = self.virtual_polys[f].sub(VirtualPolyExpr.Input(g), minus_one)

    def _tables(self):
        ptrs, flags = zip(*[_table_ptr(t) for t in self.polynomials]) if self.polynomials else ((), ())
        assert all(flags) or not any(flags), "tables must be all on the host or all on the device"
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        return arr, len(ptrs), int(bool(flags and flags[0]))


@dataclass
class SumcheckProof:
    """hyperplonk/src/piops/sumcheck.rs:15-19.  r_polys[j] = coefficient array (len_j, 32), trailing zeros trimmed."""
    num_vars: int
    claimed_sum: np.ndarray
    r_polys: List[np.ndarray]

    @staticmethod
    def prove(ctx: Context, num_vars: int, store: VirtualPolynomialStore, h: int, claimed_sum: np.ndarray,
              transcript: Transcript, sharded: bool = False):
        """sumcheck.rs:28-114 -> (SumcheckProof, EvaluationClaim).  `transcript.state` is advanced in place."""
        nd, cs = store.virtual_polys[h].flatten()
        tabs, k, on_dev = store._tables()
        mc = _lib.QZ_MAX_ROUND_COEFFS
        coeffs = np.zeros((max(num_vars, 1), mc, 32), dtype=np.uint8)
        lens = np.zeros(max(num_vars, 1), dtype=np.uint32)
        point = np.zeros((max(num_vars, 1), 32), dtype=np.uint8)
        ev = np.zeros(32, dtype=np.uint8)
        cl = _u8(claimed_sum, (32,))
        fn = ctx.lib.qz_sumcheck_prove_sharded if sharded else ctx.lib.qz_sumcheck_prove
        ctx.check(fn(ctx.h, num_vars, k, tabs, on_dev, _ptr(nd), nd.shape[0], _ptr(cs) if cs.shape[0] else None,
                     cs.shape[0], _ptr(cl), _ptr(transcript.state), mc, _ptr(coeffs), _ptr(lens), _ptr(point), _ptr(ev)))
        polys = [coeffs[j, : lens[j]].copy() for j in range(num_vars)]
        return SumcheckProof(num_vars, cl.copy(), polys), EvaluationClaim(point[:num_vars].copy(), ev)


@dataclass
class ZeroCheckProof:
    """hyperplonk/src/piops/zerocheck.rs:9-12."""
    num_vars: int
    sumcheck_proof: SumcheckProof
    z: np.ndarray  # the eq challenges (not part of the reference struct; exposed for checking)

    @staticmethod
    def prove(ctx: Context, store: VirtualPolynomialStore, h: int, transcript: Transcript, sharded: bool = False):
        """zerocheck.rs:14-49 -> (ZeroCheckProof, EvaluationClaim).  The eq table is built on the device.
        sharded: the store holds this rank's shard of every table (see SumcheckProof.prove)."""
        num_vars = store.num_vars
        nd, cs = store.virtual_polys[h].flatten()
        tabs, k, on_dev = store._tables()
        mc = _lib.QZ_MAX_ROUND_COEFFS
        coeffs = np.zeros((max(num_vars, 1), mc, 32), dtype=np.uint8)
        lens = np.zeros(max(num_vars, 1), dtype=np.uint32)
        point = np.zeros((max(num_vars, 1), 32), dtype=np.uint8)
        z = np.zeros((max(num_vars, 1), 32), dtype=np.uint8)
        ev = np.zeros(32, dtype=np.uint8)
        fn = ctx.lib.qz_zerocheck_prove_sharded if sharded else ctx.lib.qz_zerocheck_prove
        ctx.check(fn(ctx.h, num_vars, k, tabs, on_dev, _ptr(nd), nd.shape[0],
                     _ptr(cs) if cs.shape[0] else None, cs.shape[0], _ptr(transcript.state),
                     mc, _ptr(coeffs), _ptr(lens), _ptr(point), _ptr(ev), _ptr(z)))
        polys = [coeffs[j, : lens[j]].copy() for j in range(num_vars)]
        sc = SumcheckProof(num_vars, np.zeros(32, dtype=np.uint8), polys)
        return ZeroCheckProof(num_vars, sc, z[:num_vars].copy()), EvaluationClaim(point[:num_vars].copy(), ev)


def fast_eq_eval_hypercube(ctx: Context, n: int, point: np.ndarray, device: bool = False):
    """hyperplonk/src/utils/eq_eval.rs:6-31 -> (2^n, 32) host array, or a DeviceBuffer when device=True."""
    point = _u8(point, (-1, 32))
    assert point.shape[0] == n
    if device:
        buf = ctx.alloc(32 << n)
        ctx.check(ctx.lib.qz_eq_table(ctx.h, n, _ptr(point) if n else None, buf.ptr, 1))
        return buf
    out = np.zeros((1 << n, 32), dtype=np.uint8)
    ctx.check(ctx.lib.qz_eq_table(ctx.h, n, _ptr(point) if n else None, _ptr(out), 0))
    return out


def logup_denominators(ctx: Context, store: VirtualPolynomialStore, h: int, gamma: np.ndarray,
                       multiplicities: Optional[int] = None, device: bool = False):
    """hyperplonk/src/piops/multiset_check.rs:43-95: out[i] = m(row_i) / (gamma + h(row_i)); raises like the reference's
    `.inverse().unwrap()` when a denominator is zero."""
    nh, ch = store.virtual_polys[h].flatten()
    consts = ch
    nm = np.zeros((0, 3), dtype=np.uint32)
    if multiplicities is not None:
        nm, cm = store.virtual_polys[multiplicities].flatten()
        nm = nm.copy()
        nm[nm[:, 0] == 1, 1] += ch.shape[0]  # Const indices of m follow h's in the shared consts array
        consts = np.concatenate([ch, cm]) if cm.shape[0] else ch
    tabs, k, on_dev = store._tables()
    out = ctx.alloc(32 << store.num_vars) if device else np.zeros((1 << store.num_vars, 32), dtype=np.uint8)
    rc = ctx.lib.qz_logup_denominators(ctx.h, store.num_vars, k, tabs, on_dev, _ptr(nh), nh.shape[0],
                                       _ptr(np.ascontiguousarray(nm)) if nm.shape[0] else None, nm.shape[0],
                                       _ptr(np.ascontiguousarray(consts)) if consts.shape[0] else None, consts.shape[0],
                                       _ptr(_u8(gamma, (32,))), out.ptr if device else _ptr(out), int(device))
    if rc == _lib.QZ_ERR_INVALID_ARG and b"zero" in ctx.lib.qz_last_error(ctx.h):
        raise ZeroDivisionError("called `Option::unwrap()` on a `None` value (inverse of zero)")
    ctx.check(rc)
    return out
