"""Run ONE call of a hot path (for ncu): python tools/profile_one.py sumcheck|msm|zerocheck|mlpcs LOG_N"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import quill_zkvm_b200 as q

FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def mont(v, mod=FR):
    return np.frombuffer(((v % mod) * (1 << 256) % mod).to_bytes(32, "little"), dtype=np.uint8).copy()


what, n = sys.argv[1], int(sys.argv[2])
ctx = q.Context(0)
if what in ("sumcheck", "zerocheck"):
    store = q.VirtualPolynomialStore(n)
    for t in range(3):
        store.allocate_polynomial(ctx.random_fr(1 << n, 5 + t))
    e = q.VirtualPolyExpr.Input(0) * q.VirtualPolyExpr.Input(1) * q.VirtualPolyExpr.Input(2)
    h = store.new_virtual_from_expr(e)
    for rep in range(2):
        tr = q.Transcript(b"profile", ctx)
        if what == "sumcheck":
            q.SumcheckProof.prove(ctx, n, store, h, mont(1), tr)
        else:
            q.ZeroCheckProof.prove(ctx, store, h, tr)
        print(what, n, "ms", ctx.last_elapsed_ms(0), "rounds ms", ctx.last_elapsed_ms(1))
elif what == "mlpcs":
    g = np.concatenate([mont(1, FQ), mont(2, FQ)])
    kzg = q.KZG.trusted_setup(ctx, (1 << n) - 1, g, mont(0x1234567)).precompute()
    poly = ctx.random_fr(1 << n, 777)
    point = np.frombuffer(b"".join(((i * 0x9E3779B97F4A7C15 + 12345) % FR).to_bytes(32, "little") for i in range(n)),
                          dtype=np.uint8).reshape(n, 32).copy()
    for rep in range(2):
        kzg.commit(poly)
        kzg.open_multilinear(poly, point, q.Transcript(b"mlpcs_bench", ctx))
        print("mlpcs", n, "open ms", ctx.last_elapsed_ms(0))
else:
    g = np.concatenate([mont(1, FQ), mont(2, FQ)])
    kzg = q.KZG.trusted_setup(ctx, (1 << n) - 1, g, mont(0x1234567))
    if len(sys.argv) > 3 and sys.argv[3] == "pre":
        kzg.precompute()
    sc = ctx.random_fr(1 << n, 9)
    for rep in range(2):
        kzg.commit(sc)
        print("msm", n, "ms", ctx.last_elapsed_ms(0), "accumulate ms", ctx.last_elapsed_ms(1))
ctx.close()
