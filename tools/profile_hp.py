"""One HyperPlonk proof (Fibonacci, 2^K rows) for profiling: python tools/profile_hp.py K"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quill_zkvm_b200 as q
from quill_zkvm_b200 import hyperplonk as hp
import bench

K = int(sys.argv[1])
ctx = q.Context(0)
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
mont = lambda v, mod=bench.FR: np.frombuffer(((v % mod) * (1 << 256) % mod).to_bytes(32, "little"), dtype=np.uint8).copy()
g = np.concatenate([mont(1, FQ), mont(2, FQ)])
def timed_loop(fn, steps, warmup):
    for _ in range(warmup): fn()
    ctx.sync(); t=time.perf_counter(); l0=ctx.kernel_launches
    for _ in range(steps): fn()
    ctx.sync(); return (time.perf_counter()-t)*1e3/steps, ctx.kernel_launches-l0
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 1
def timed_loop_reps(fn, steps, warmup):
    for _ in range(warmup): fn()
    out = []
    for _ in range(REPS):
        ctx.sync(); t=time.perf_counter(); l0=ctx.kernel_launches
        fn()
        ctx.sync(); out.append(round((time.perf_counter()-t)*1e3, 1))
    print("per-proof ms:", out)
    return min(out), ctx.kernel_launches-l0
r = bench.bench_hyperplonk(ctx, q, K, g, mont(bench.TAU), timed_loop_reps if REPS > 1 else timed_loop)
print({k: r[k] for k in ("value", "rows_per_trace", "gpu_launches")})
