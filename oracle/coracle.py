"""ctypes binding of oracle/liboracle.so (CPU ORACLE -- test infrastructure, not product code).

Values cross this boundary as numpy uint8 arrays: Fr/Fq = (..., 32) little-endian Montgomery limbs,
G1 affine = (..., 64) x‖y Montgomery with all-zero = infinity.  Helpers convert to/from Python ints.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R256 = 1 << 256
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("capi.cpp", "field.hpp", "g1.hpp", "pairing.hpp", "blake3_ref.hpp", "protocol.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_sumcheck_prove.restype = C.c_int
        _LIB.orc_sumcheck_verify.restype = C.c_int
        _LIB.orc_g1_on_curve.restype = C.c_int
        _LIB.orc_kzg_commit_reference_shape.restype = C.c_int
        _LIB.orc_mlpcs_open.restype = C.c_int
        _LIB.orc_logup_denominators.restype = C.c_int
        _LIB.orc_g2_on_curve.restype = C.c_int
        _LIB.orc_pairing_product.restype = C.c_int
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


# ---- int <-> Montgomery bytes ---------------------------------------------------------------------------------
def to_mont(vals, mod=FR) -> np.ndarray:
    vals = list(vals)
    out = np.zeros((len(vals), 32), dtype=np.uint8)
    for i, v in enumerate(vals):
        out[i] = np.frombuffer(((v % mod) * R256 % mod).to_bytes(32, "little"), dtype=np.uint8)
    return out


def from_mont(arr: np.ndarray, mod=FR):
    arr = np.ascontiguousarray(arr, dtype=np.uint8).reshape(-1, 32)
    rinv = pow(R256, -1, mod)
    return [int.from_bytes(arr[i].tobytes(), "little") * rinv % mod for i in range(arr.shape[0])]


def fr1(v) -> np.ndarray:
    return to_mont([v])[0]


def g1_to_bytes(p) -> np.ndarray:
    if p is None:
        return np.zeros(64, dtype=np.uint8)
    return np.concatenate([to_mont([p[0]], FQ)[0], to_mont([p[1]], FQ)[0]])


def g1_from_bytes(b: np.ndarray):
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(64)
    if not b.any():
        return None
    x, = from_mont(b[:32], FQ)
    y, = from_mont(b[32:], FQ)
    return (x, y)


# ---- wrappers -----------------------------------------------------------------------------------------------------
def blake3(data: bytes, out_len: int = 32) -> bytes:
    out = np.zeros(out_len, dtype=np.uint8)
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if data else np.zeros(0, dtype=np.uint8)
    lib().orc_blake3(_p(np.ascontiguousarray(buf)), C.c_size_t(len(data)), _p(out), C.c_size_t(out_len))
    return out.tobytes()


def transcript_new(domain: bytes) -> np.ndarray:
    st = np.zeros(32, dtype=np.uint8)
    d = np.frombuffer(bytes(domain), dtype=np.uint8).copy() if domain else np.zeros(0, dtype=np.uint8)
    lib().orc_transcript_new(_p(d), C.c_size_t(len(domain)), _p(st))
    return st


def transcript_append(state: np.ndarray, msg: bytes):
    m = np.frombuffer(bytes(msg), dtype=np.uint8).copy() if msg else np.zeros(0, dtype=np.uint8)
    lib().orc_transcript_append(_p(state), _p(m), C.c_size_t(len(msg)))


def transcript_draw_fr(state: np.ndarray) -> np.ndarray:
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_transcript_draw_fr(_p(state), _p(out))
    return out


def field_op(field: int, op: int, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros_like(a)
    bp = _p(np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 32)) if b is not None else None
    lib().orc_field_op(C.c_int(field), C.c_int(op), _p(a), bp, _p(out), C.c_size_t(a.shape[0]))
    return out


def g1_add(a, b):
    out = np.zeros(64, dtype=np.uint8)
    lib().orc_g1_add(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
    return out


def g1_mul(a, scalar_mont):
    out = np.zeros(64, dtype=np.uint8)
    lib().orc_g1_mul(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(scalar_mont)), _p(out))
    return out


def g1_on_curve(a) -> bool:
    return bool(lib().orc_g1_on_curve(_p(np.ascontiguousarray(a))))


def g1_serialize(a) -> bytes:
    out = np.zeros(64, dtype=np.uint8)
    lib().orc_g1_serialize(_p(np.ascontiguousarray(a)), _p(out))
    return out.tobytes()


def srs_generate(g_xy: np.ndarray, tau_mont: np.ndarray, n: int, threads: int = 0) -> np.ndarray:
    out = np.zeros((n, 64), dtype=np.uint8)
    lib().orc_srs_generate(_p(np.ascontiguousarray(g_xy)), _p(np.ascontiguousarray(tau_mont)), C.c_size_t(n), _p(out),
                           C.c_int(threads or os.cpu_count() or 1))
    return out


def msm(bases: np.ndarray, scalars: np.ndarray, mode: int = 1, threads: int = 1) -> np.ndarray:
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros(64, dtype=np.uint8)
    lib().orc_msm(_p(bases), C.c_size_t(bases.shape[0]), _p(scalars), C.c_size_t(scalars.shape[0]), C.c_int(mode),
                  C.c_int(threads), _p(out))
    return out


def kzg_commit_reference_shape(bases: np.ndarray, scalars: np.ndarray, threads: int = 1):
    """KZG::commit as pcs/src/kzg.rs:61-73 runs it (per-call SRS normalisation + Pippenger).
    Returns (point bytes, seconds_normalise, seconds_msm)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros(64, dtype=np.uint8)
    secs = (C.c_double * 2)()
    rc = lib().orc_kzg_commit_reference_shape(_p(bases), C.c_size_t(bases.shape[0]), _p(scalars),
                                              C.c_size_t(scalars.shape[0]), C.c_int(threads), _p(out), secs)
    if rc:
        raise AssertionError("Polynomial degree exceeds max degree")
    return out, secs[0], secs[1]


def kzg_open_quotient(poly: np.ndarray, x_mont: np.ndarray):
    poly = np.ascontiguousarray(poly, dtype=np.uint8).reshape(-1, 32)
    n = poly.shape[0]
    y = np.zeros(32, dtype=np.uint8)
    q = np.zeros((max(n - 1, 0), 32), dtype=np.uint8)
    qlen = C.c_size_t(0)
    lib().orc_kzg_open_quotient(_p(poly), C.c_size_t(n), _p(np.ascontiguousarray(x_mont)), _p(y), _p(q), C.byref(qlen))
    return y, q[: qlen.value]


def eq_table(point: np.ndarray) -> np.ndarray:
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    n = point.shape[0]
    out = np.zeros((1 << n, 32), dtype=np.uint8)
    lib().orc_eq_table(C.c_size_t(n), _p(point), _p(out))
    return out


def eq_eval(x: np.ndarray, r: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.uint8).reshape(-1, 32)
    r = np.ascontiguousarray(r, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_eq_eval(C.c_size_t(x.shape[0]), _p(x), _p(r), _p(out))
    return out


def mle_evaluate(evals: np.ndarray, point: np.ndarray) -> np.ndarray:
    evals = np.ascontiguousarray(evals, dtype=np.uint8).reshape(-1, 32)
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    assert evals.shape[0] == 1 << point.shape[0]
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_mle_evaluate(C.c_size_t(point.shape[0]), _p(evals), _p(point), _p(out))
    return out


def _expr_arrays(nodes, consts):
    nd = np.ascontiguousarray(np.array(nodes, dtype=np.uint32).reshape(-1, 3))
    cs = np.ascontiguousarray(np.array(consts, dtype=np.uint8).reshape(-1, 32)) if len(consts) else np.zeros((0, 32), np.uint8)
    return nd, cs


def sumcheck_prove(num_vars, tables, nodes, consts, claimed_sum, state, max_coeffs=8, zerocheck=False, threads=1):
    """tables: list of (2^n, 32) uint8 Montgomery arrays; nodes: [(op,a,b)...]; consts: (m,32) Montgomery.
    Returns dict(coeffs (n,max_coeffs,32), lens (n,), point (n,32), evaluation (32,), z (n,32)|None); state is updated in place."""
    tabs = [np.ascontiguousarray(t, dtype=np.uint8).reshape(-1, 32) for t in tables]
    for t in tabs:
        assert t.shape[0] == 1 << num_vars
    ptrs = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    nd, cs = _expr_arrays(nodes, consts)
    coeffs = np.zeros((num_vars, max_coeffs, 32), dtype=np.uint8)
    lens = np.zeros(num_vars, dtype=np.uint32)
    point = np.zeros((num_vars, 32), dtype=np.uint8)
    ev = np.zeros(32, dtype=np.uint8)
    z = np.zeros((num_vars, 32), dtype=np.uint8)
    cl = np.ascontiguousarray(claimed_sum, dtype=np.uint8) if claimed_sum is not None else np.zeros(32, np.uint8)
    rc = lib().orc_sumcheck_prove(C.c_size_t(num_vars), C.c_size_t(len(tabs)), ptrs, _p(nd), C.c_size_t(nd.shape[0]),
                                  _p(cs), C.c_size_t(cs.shape[0]), _p(cl), _p(state), C.c_size_t(max_coeffs),
                                  _p(coeffs), _p(lens), _p(point), _p(ev), C.c_int(1 if zerocheck else 0), _p(z),
                                  C.c_int(threads))
    if rc:
        raise RuntimeError(f"oracle sumcheck_prove rc={rc}")
    return dict(coeffs=coeffs, lens=lens, point=point, evaluation=ev, z=z if zerocheck else None)


def sumcheck_verify(num_vars, claimed_sum, coeffs, lens, state):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint8)
    max_coeffs = coeffs.shape[1] if num_vars else 1
    point = np.zeros((num_vars, 32), dtype=np.uint8)
    ev = np.zeros(32, dtype=np.uint8)
    rc = lib().orc_sumcheck_verify(C.c_size_t(num_vars), _p(np.ascontiguousarray(claimed_sum)), C.c_size_t(max_coeffs),
                                   _p(coeffs), _p(np.ascontiguousarray(lens, dtype=np.uint32)), _p(state), _p(point),
                                   _p(ev))
    return (rc == 0), point, ev


def expr_eval_point(nodes, consts, g: np.ndarray) -> np.ndarray:
    nd, cs = _expr_arrays(nodes, consts)
    g = np.ascontiguousarray(g, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_expr_eval_point(_p(nd), C.c_size_t(nd.shape[0]), _p(cs), C.c_size_t(cs.shape[0]), _p(g),
                              C.c_size_t(g.shape[0]), _p(out))
    return out


def compute_s_polynomial(p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    p1 = np.ascontiguousarray(p1, dtype=np.uint8).reshape(-1, 32)
    p2 = np.ascontiguousarray(p2, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros((max(p1.shape[0], p2.shape[0], 1), 32), dtype=np.uint8)
    n = C.c_size_t(0)
    lib().orc_compute_s_polynomial(_p(p1), C.c_size_t(p1.shape[0]), _p(p2), C.c_size_t(p2.shape[0]), _p(out), C.byref(n))
    return out[: n.value]


def compute_pr(point: np.ndarray) -> np.ndarray:
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros((1 << point.shape[0], 32), dtype=np.uint8)
    n = C.c_size_t(0)
    lib().orc_compute_pr(_p(point), C.c_size_t(point.shape[0]), _p(out), C.byref(n))
    return out[: n.value]


def mlpcs_open(bases: np.ndarray, poly: np.ndarray, point: np.ndarray, state: np.ndarray, threads: int = 1):
    """MLEvalProof::prove (pcs/src/mlpcs.rs:83-124).  Returns dict(evaluation, s_comm, openings=[(x, y, proof) x 4])."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    poly = np.ascontiguousarray(poly, dtype=np.uint8).reshape(-1, 32)
    point = np.ascontiguousarray(point, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros(608, dtype=np.uint8)
    rc = lib().orc_mlpcs_open(_p(bases), C.c_size_t(bases.shape[0]), _p(poly), C.c_size_t(poly.shape[0]), _p(point),
                              C.c_size_t(point.shape[0]), _p(state), C.c_int(threads), _p(out))
    if rc:
        raise AssertionError("Polynomial degree exceeds max degree")
    ops = [(out[96 + 128 * i: 128 + 128 * i].copy(), out[128 + 128 * i: 160 + 128 * i].copy(),
            out[160 + 128 * i: 224 + 128 * i].copy()) for i in range(4)]
    return dict(evaluation=out[:32].copy(), s_comm=out[32:96].copy(), openings=ops)


def logup_denominators(num_vars, tables, nodes_h, nodes_m, consts, gamma):
    """multiset_check.rs:43-95.  nodes_m = [] for Equality mode.  Raises ZeroDivisionError like the reference's unwrap."""
    tabs = [np.ascontiguousarray(t, dtype=np.uint8).reshape(-1, 32) for t in tables]
    ptrs = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    nh, cs = _expr_arrays(nodes_h, consts)
    nm = np.ascontiguousarray(np.array(nodes_m, dtype=np.uint32).reshape(-1, 3))
    out = np.zeros((1 << num_vars, 32), dtype=np.uint8)
    rc = lib().orc_logup_denominators(C.c_size_t(num_vars), C.c_size_t(len(tabs)), ptrs, _p(nh), C.c_size_t(nh.shape[0]),
                                      _p(nm), C.c_size_t(nm.shape[0]), _p(cs), C.c_size_t(cs.shape[0]),
                                      _p(np.ascontiguousarray(gamma, dtype=np.uint8)), _p(out))
    if rc:
        raise ZeroDivisionError("inverse of zero")
    return out


# ---- verifier side: G2 and the pairing (pcs/src/kzg.rs:49-52, 98-108) ---------------------------------------------
# G2 affine = (128,) bytes x.c0 ‖ x.c1 ‖ y.c0 ‖ y.c1 (Montgomery Fq), all-zero = infinity.
FINAL_EXP = ((FQ ** 12 - 1) // FR).to_bytes(((FQ ** 12 - 1) // FR).bit_length() // 8 + 1, "little")


def g2_generator() -> np.ndarray:
    out = np.zeros(128, dtype=np.uint8)
    lib().orc_g2_generator(_p(out))
    return out


def g2_on_curve(a) -> bool:
    return bool(lib().orc_g2_on_curve(_p(np.ascontiguousarray(a))))


def g2_add(a, b) -> np.ndarray:
    out = np.zeros(128, dtype=np.uint8)
    lib().orc_g2_add(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
    return out


def g2_neg(a) -> np.ndarray:
    out = np.zeros(128, dtype=np.uint8)
    lib().orc_g2_neg(_p(np.ascontiguousarray(a)), _p(out))
    return out


def g2_mul(a, scalar_mont) -> np.ndarray:
    out = np.zeros(128, dtype=np.uint8)
    lib().orc_g2_mul(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(scalar_mont)), _p(out))
    return out


def pairing_product(g1s, g2s):
    """prod_i e(P_i, Q_i): (is_one, the Fq12 value as 12 canonical ints)."""
    g1s = np.ascontiguousarray(np.asarray(g1s, dtype=np.uint8).reshape(-1, 64))
    g2s = np.ascontiguousarray(np.asarray(g2s, dtype=np.uint8).reshape(-1, 128))
    assert g1s.shape[0] == g2s.shape[0]
    e = np.frombuffer(FINAL_EXP, dtype=np.uint8).copy()
    out = np.zeros((12, 32), dtype=np.uint8)
    one = lib().orc_pairing_product(_p(g1s), _p(g2s), C.c_size_t(g1s.shape[0]), _p(e), C.c_size_t(e.shape[0]), _p(out))
    return bool(one), from_mont(out, FQ)
