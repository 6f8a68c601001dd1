"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/quill_b200.h declares,
refuses to run without a device (no CPU fallback), and its host-side transcript matches the oracle."""
import os
import re

import numpy as np
import pytest

import quill_zkvm_b200 as q
from oracle import coracle as co
from quill_zkvm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "quill_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(qz_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 35
    assert sorted(_lib.SYMBOLS) == declared, "python binding list and header disagree"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libquill_b200.so"


def test_header_compiles_as_c_and_links(tmp_path):
    """include/quill_b200.h is plain C (what bindgen / cgo / a C caller would consume): a C99 translation unit that
    includes nothing else compiles without warnings, links against the shared library and gets the documented refusal
    when there is no device."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "abi_link.c"
    src.write_text(
        '#include "quill_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  qz_ctx* ctx = 0;\n"
        "  int rc = qz_ctx_create(0, 0, &ctx);\n"
        '  printf("%d %s\\n", rc, qz_status_str(rc));\n'
        "  if (ctx) qz_ctx_destroy(ctx);\n"
        "  return 0;\n}\n")
    exe = tmp_path / "abi_link"
    libdir = os.path.join(ROOT, "quill_zkvm_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lquill_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    rc = int(out.stdout.split()[0])
    assert rc == (_lib.QZ_OK if _has_gpu() else _lib.QZ_ERR_NO_DEVICE), out.stdout


def test_no_cpu_fallback():
    if _has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(q.QuillError) as e:
        q.Context(0)
    assert e.value.status == _lib.QZ_ERR_NO_DEVICE


def test_status_strings():
    lib = _lib.load()
    assert lib.qz_status_str(0) == b"ok"
    assert b"degree" in lib.qz_status_str(_lib.QZ_ERR_DEGREE).lower()


@pytest.mark.parametrize("domain", [b"", b"sumcheck_test", b"hyperplonk_proof", b"q" * 3000])
def test_host_transcript_matches_oracle(domain):
    t = q.Transcript(domain)
    st = co.transcript_new(domain)
    assert t.state.tobytes() == st.tobytes()
    for msg in (b"", b"abc", bytes(range(256)) * 5, b"\x00" * 1024):
        t.append_bytes(msg)
        co.transcript_append(st, msg)
        assert t.state.tobytes() == st.tobytes()
    for n in (1, 32, 48, 64, 65, 300):
        c = t.draw_challenge(n)
        # oracle: blake3-XOF(state ‖ "challenge") then re-absorb
        want = co.blake3(st.tobytes() + b"challenge", n)
        co.transcript_append(st, want)
        assert c == want and t.state.tobytes() == st.tobytes()


def test_expr_flatten_roundtrip():
    m1 = co.fr1(co.FR - 1)
    e = (q.VirtualPolyExpr.Input(0) * q.VirtualPolyExpr.Input(0)).sub(q.VirtualPolyExpr.Input(1), m1)
    nodes, consts = e.flatten()
    assert nodes.tolist() == [[0, 0, 0], [0, 0, 0], [3, 0, 1], [1, 0, 0], [0, 1, 0], [3, 3, 4], [2, 2, 5]]
    assert consts.shape == (1, 32)
    g = co.to_mont([6, 10])
    assert co.from_mont(co.expr_eval_point([tuple(r) for r in nodes.tolist()], consts, g))[0] == 26


def test_rust_sys_crate_declares_every_symbol():
    """rust/quill-b200-sys/src/lib.rs (uncompiled: no Rust toolchain here) must at least name every function the header
    declares, with the same number of parameters"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "quill_b200.h")).read()
    rust = open(os.path.join(root, "rust", "quill-b200-sys", "src", "lib.rs")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    c_fns = {m.group(1): m.group(2) for m in re.finditer(r"\b(qz_\w+)\s*\(([^;{]*?)\)\s*;", header)}
    r_fns = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (qz_\w+)\s*\(([^;]*?)\)\s*(?:->[^;]*)?;", rust, flags=re.S)}
    assert len(c_fns) >= 50
    missing = sorted(set(c_fns) - set(r_fns))
    assert not missing, missing

    def arity(params: str) -> int:
        params = params.strip()
        return 0 if params in ("", "void") else params.count(",") + 1
    wrong = [f for f in c_fns if arity(c_fns[f]) != arity(r_fns[f])]
    assert not wrong, wrong
