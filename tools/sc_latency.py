"""Per-round latency of the sumcheck prover: proofs over 2^n-entry tables for small n are pure latency (n <= 11: the
single-block tail only; every further variable adds one streaming round = round kernel + finalize)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quill_zkvm_b200 as q  # noqa: E402
from bench import product_expr  # noqa: E402


def main():
    stream = torch.cuda.Stream()
    ctx = q.Context(0, stream.cuda_stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    claimed = np.zeros(32, np.uint8)
    prev = None
    for nv in (4, 8, 11, 12, 13, 14, 16, 18, 20, 22, 24):
        tabs = [ctx.random_fr(1 << nv, 5 + t) for t in range(3)]
        store = q.VirtualPolynomialStore(nv)
        store.polynomials = tabs
        store.virtual_polys = [product_expr(q, 3)]

        def prove():
            tr = q.Transcript(b"lat", ctx)
            q.SumcheckProof.prove(ctx, nv, store, 0, claimed, tr)

        with torch.cuda.stream(stream):
            for _ in range(3):
                prove()
            torch.cuda.synchronize()
            ev0.record(stream)
            reps = 10
            l0 = ctx.kernel_launches
            dev = []
            for _ in range(reps):
                prove()
                dev.append(ctx.last_elapsed_ms(0))
            ev1.record(stream)
            torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / reps
        # "on the stream": events recorded by the library around its own enqueues (input copy .. result copy), i.e. without
        # the host's preparation between two calls, which depends on the box's CPU
        print(f"n={nv:2d}  {ms * 1e3:9.1f} us/proof  on the stream {min(dev) * 1e3:8.1f} us  {ms * 1e3 / nv:7.1f} us/round  launches {(ctx.kernel_launches - l0) // reps}"
              + (f"  delta vs previous size {1e3 * (ms - prev[1]) / (nv - prev[0]):7.1f} us/added round" if prev else ""), flush=True)
        prev = (nv, ms)
        for t in tabs:
            t.free()
    ctx.close()


if __name__ == "__main__":
    main()
