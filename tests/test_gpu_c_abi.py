"""The C ABI driven from a plain C program (tests/c/abi_golden.c): qz_sumcheck_prove and qz_kzg_commit / qz_kzg_open on the
reference's 3-variable / degree-2 test shapes, compared with tests/golden/golden.json -- no Python between the caller
and the library.  The expected values are written into a generated header; gcc builds the program against
include/quill_b200.h and the in-tree libquill_b200.so."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def _le(hexstr: str) -> bytes:  # "%064x" -> 32 little-endian bytes
    return int(hexstr, 16).to_bytes(32, "little")


def _arr(name: str, data: bytes) -> str:
    return f"static const uint8_t {name}[{len(data)}] = {{{','.join(str(b) for b in data)}}};\n"


def write_header(path: str):
    sc, kz = GOLDEN["sumcheck_test"], GOLDEN["kzg_test"]
    out = "#include <stdint.h>\n"
    out += _arr("GOLDEN_SUMCHECK_STATE_END", bytes.fromhex(sc["state_end"]))
    out += f"static const uint32_t GOLDEN_SUMCHECK_LENS[3] = {{{','.join(str(len(p)) for p in sc['r_polys'])}}};\n"
    width = 33 * 32
    rows = [b"".join(_le(c) for c in p).ljust(width, b"\0") for p in sc["r_polys"]]
    out += "static const uint8_t GOLDEN_SUMCHECK_RPOLYS[3][%d] = {%s};\n" % (
        width, ",".join("{" + ",".join(str(b) for b in r) + "}" for r in rows))
    out += _arr("GOLDEN_SUMCHECK_POINT", b"".join(_le(x) for x in sc["point"]))
    out += _arr("GOLDEN_SUMCHECK_EVALUATION", _le(sc["evaluation"]))
    out += _arr("GOLDEN_KZG_G", _le(kz["g"][0]) + _le(kz["g"][1]))
    out += _arr("GOLDEN_KZG_TAU", _le(kz["tau"]))
    out += _arr("GOLDEN_KZG_COMMITMENT_BYTES", bytes.fromhex(kz["commitment_bytes"]))
    out += _arr("GOLDEN_KZG_Y", _le(kz["y"]))
    out += _arr("GOLDEN_KZG_PROOF", _le(kz["proof"][0]) + _le(kz["proof"][1]))
    open(path, "w").write(out)


def build(tmp: str) -> str:
    write_header(os.path.join(tmp, "golden_c.h"))
    exe = os.path.join(tmp, "abi_golden")
    libdir = os.path.join(ROOT, "quill_zkvm_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", tmp, "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_golden.c"), "-o", exe, "-L", libdir, "-lquill_b200",
                           f"-Wl,-rpath,{libdir}"])
    return exe


def test_c_program_builds(tmp_path):
    """CPU: the header generator and the C caller compile and link against the library (no compute without a GPU: the
    program must stop at qz_ctx_create with QZ_ERR_NO_DEVICE)"""
    exe = build(str(tmp_path))
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run is the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c_program_matches_golden(tmp_path):
    exe = build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C ABI golden check ok" in r.stdout
