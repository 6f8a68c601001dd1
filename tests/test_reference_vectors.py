"""Pins the oracle (and through tests/golden/golden.json, which every GPU parity test holds the CUDA path to) to the REAL
reference when its vectors are present.

tests/golden/reference.json is written by tools/refvec -- a small Rust program that runs the unmodified reference
(arkworks 0.5.0, blake3 1.8.2) on golden.json's inputs.  This image has no Rust toolchain, so the file cannot be
produced here ("parity unpinned", oracle/README.md); on a machine with cargo:

    cargo run --release --manifest-path tools/refvec/Cargo.toml -- tests/golden/golden.json tests/golden/reference.json

When the file exists these tests compare it key for key with golden.json (made by oracle/pyref.py) and re-derive the
MLPCS opening with the C++ oracle; when it does not, they skip and say why.  A second, always-on test keeps golden.json
honest against both restatements for the section the emitter added (mlpcs_n5).
"""
import json
import os

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyref as py

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))
REF_PATH = os.path.join(HERE, "golden", "reference.json")
FR = py.FR
GEN = py.g1_mul(py.G1_GEN, 7)
TAU = 0x1234567890ABCDEF1234567890ABCDEF


def _leaves(o, path=""):
    if isinstance(o, dict):
        for k, v in o.items():
            yield from _leaves(v, f"{path}/{k}")
    elif isinstance(o, list) and o and isinstance(o[0], (dict, list)):
        for i, v in enumerate(o):
            yield from _leaves(v, f"{path}[{i}]")
    else:
        yield path, o


def _get(o, path):
    for part in path.strip("/").replace("[", "/[").split("/"):
        o = o[int(part[1:-1])] if part.startswith("[") else o[part]
    return o


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="tests/golden/reference.json absent: run tools/refvec with a Rust "
                    "toolchain (none in this image) to pin the oracle to the arkworks binary")
def test_reference_vectors_equal_golden():
    ref = json.load(open(REF_PATH))
    checked, missing = 0, []
    for path, want in _leaves({k: v for k, v in ref.items() if not k.startswith("_")}):
        try:
            got = _get(GOLDEN, path)
        except (KeyError, IndexError):
            missing.append(path)
            continue
        assert got == want, f"{path}: golden.json (oracle) {got!r} != reference.json (arkworks) {want!r}"
        checked += 1
    assert checked > 100, (checked, missing)
    # every golden section the emitter covers must actually have been compared
    for section in ("transcript", "sumcheck_test", "zerocheck_test", "zerocheck_test_not_zero", "product3_n6", "mixed_n6",
                    "eq_n5", "kzg_test", "msm64", "pr", "s_poly", "mlpcs_n5", "hyperplonk"):
        assert section in ref, section


def _srs(n):
    return co.srs_generate(co.g1_to_bytes(GEN), co.fr1(TAU), n, threads=os.cpu_count() or 1)


def test_golden_mlpcs_opening_both_oracles():
    """golden.json's mlpcs_n5 (made by pyref) re-derived by the C++ oracle: MLEvalProof::prove, mlpcs.rs:83-124"""
    m = GOLDEN["mlpcs_n5"]
    poly = co.to_mont([int(x, 16) for x in m["poly"]])
    point = co.to_mont([int(x, 16) for x in m["point"]])
    srs = _srs(65)
    com = co.msm(srs[:32], poly, mode=1)
    assert co.g1_serialize(com).hex() == m["commitment_bytes"]
    st = co.transcript_new(b"mlpcs_golden")
    got = co.mlpcs_open(srs, poly, point, st)
    assert "%064x" % co.from_mont(got["evaluation"])[0] == m["evaluation"]
    assert co.g1_serialize(got["s_comm"]).hex() == m["s_comm_bytes"]
    for (x, y, proof), key in zip(got["openings"], ("poly_opening", "poly_opening_inv", "s_opening", "s_opening_inv")):
        assert "%064x" % co.from_mont(x)[0] == m[key]["x"] and "%064x" % co.from_mont(y)[0] == m[key]["y"]
        assert co.g1_serialize(proof).hex() == m[key]["proof_bytes"]
    assert st.tobytes().hex() == m["state_end"]


@pytest.mark.gpu
def test_golden_mlpcs_opening_gpu(ctx):
    """the same vector through the C ABI (qz_kzg_commit, qz_mlpcs_open)"""
    import quill_zkvm_b200 as q

    m = GOLDEN["mlpcs_n5"]
    poly = co.to_mont([int(x, 16) for x in m["poly"]])
    point = co.to_mont([int(x, 16) for x in m["point"]])
    kzg = q.KZG.trusted_setup(ctx, 64, co.g1_to_bytes(GEN), co.fr1(TAU))
    assert ctx.g1_serialize(kzg.commit(poly)).hex() == m["commitment_bytes"]
    tr = q.Transcript(b"mlpcs_golden", ctx)
    pf = kzg.open_multilinear(poly, point, tr)
    assert "%064x" % co.from_mont(pf.evaluation)[0] == m["evaluation"]
    assert ctx.g1_serialize(pf.s_comm).hex() == m["s_comm_bytes"]
    for o, key in zip((pf.poly_opening, pf.poly_opening_inv, pf.s_opening, pf.s_opening_inv),
                      ("poly_opening", "poly_opening_inv", "s_opening", "s_opening_inv")):
        assert "%064x" % co.from_mont(o.x)[0] == m[key]["x"] and "%064x" % co.from_mont(o.y)[0] == m[key]["y"]
        assert ctx.g1_serialize(o.proof).hex() == m[key]["proof_bytes"]
    assert tr.state.tobytes().hex() == m["state_end"]
    kzg.srs.free()
