"""Timeline of the persistent round kernel (sc_mid): run with QZ_LIB_PATH=tools/_libs/libquill_trace.so (built by
tools/build_variant.sh trace -DQZ_SC_TRACE).  Prints, per round, the time between the stamps of the closing block."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quill_zkvm_b200 as q  # noqa: E402
from quill_zkvm_b200 import _lib  # noqa: E402

NAMES = {1: "round start", 8: "tile folded", 9: "tile evaluated", 2: "pairs done", 3: "block sums / arrive", 4: "closer has sums", 10: "interpolated", 11: "message ready",
         12: "absorbed", 13: "challenge drawn", 6: "closed", 7: "released"}


def main():
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    what = sys.argv[2] if len(sys.argv) > 2 else "sumcheck"
    ctx = q.Context(0)
    lib = _lib.load()
    store = q.VirtualPolynomialStore(nv)
    for t in range(3):
        store.allocate_polynomial(ctx.random_fr(1 << nv, 5 + t))
    e = q.VirtualPolyExpr.Input(0) * q.VirtualPolyExpr.Input(1) * q.VirtualPolyExpr.Input(2)
    h = store.new_virtual_from_expr(e)
    one = np.zeros(32, np.uint8)
    buf = (ctypes.c_ulonglong * (3 * 8192))()
    for rep in range(3):
        tr = q.Transcript(b"trace", ctx)
        if what == "sumcheck":
            q.SumcheckProof.prove(ctx, nv, store, h, one, tr)
        else:
            q.ZeroCheckProof.prove(ctx, store, h, tr)
        n = lib.qz_debug_trace(buf, 8192)
    print(what, nv, "call ms", ctx.last_elapsed_ms(0), "records", n)
    recs = [(buf[3 * i] & 0xffff, buf[3 * i] >> 16, buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]
    t0 = min(r[3] for r in recs)
    # per round: the closing block's stamps in order; the others only their 1 / 2 / 3 / 7
    rnd = -1
    last = None
    for ident, blk, clk, gt in sorted(recs, key=lambda r: r[3]):
        if ident == 1 and blk == 0:
            rnd += 1
        if (blk == 0 and ident not in (8, 9)) or ident in (4, 10, 11, 12, 13, 6) or (blk == 0 and ident in (8, 9)):
            d = "" if last is None else f"+{(gt - last) / 1e3:7.2f} us"
            print(f"round {rnd:2d} blk {blk:3d} {NAMES.get(ident, ident):22s} t={(gt - t0) / 1e3:9.2f} us {d}  clk {clk}")
            last = gt
    ctx.close()


if __name__ == "__main__":
    main()
