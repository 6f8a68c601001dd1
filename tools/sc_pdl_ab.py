"""A/B of programmatic dependent launches in the sumcheck round chain: two contexts in one process (QZ_NO_PDL is read
when a context is created), alternated so that clocks and thermals are shared."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quill_zkvm_b200 as q  # noqa: E402
from bench import product_expr  # noqa: E402


def main():
    stream = torch.cuda.Stream()
    os.environ.pop("QZ_NO_PDL", None)
    ctx_a = q.Context(0, stream.cuda_stream)
    os.environ["QZ_NO_PDL"] = "1"
    ctx_b = q.Context(0, stream.cuda_stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    claimed = np.zeros(32, np.uint8)
    for nv in (14, 16, 20, 22, 24):
        res = {"pdl": [], "plain": []}
        stores = {}
        for name, ctx in (("pdl", ctx_a), ("plain", ctx_b)):
            tabs = [ctx.random_fr(1 << nv, 5 + t) for t in range(3)]
            st = q.VirtualPolynomialStore(nv)
            st.polynomials = tabs
            st.virtual_polys = [product_expr(q, 3)]
            stores[name] = (ctx, st, tabs)
        for rep in range(4):
            for name in ("pdl", "plain"):
                ctx, st, _ = stores[name]

                def prove():
                    q.SumcheckProof.prove(ctx, nv, st, 0, claimed, q.Transcript(b"lat", ctx))

                with torch.cuda.stream(stream):
                    prove()
                    torch.cuda.synchronize()
                    ev0.record(stream)
                    for _ in range(10):
                        prove()
                    ev1.record(stream)
                    torch.cuda.synchronize()
                res[name].append(ev0.elapsed_time(ev1) / 10 * 1e3)
        print(f"n={nv}: pdl {min(res['pdl']):8.1f} us (all {[round(x) for x in res['pdl']]})   plain {min(res['plain']):8.1f} us (all {[round(x) for x in res['plain']]})", flush=True)
        for name in stores:
            for t in stores[name][2]:
                t.free()
    ctx_a.close()
    ctx_b.close()


if __name__ == "__main__":
    main()
