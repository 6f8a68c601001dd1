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
# wall time per kind of library call, per proof (every call blocks until its stream work is done)
ACC = {}
def _wrap(obj, name, key=None):
    fn = getattr(obj, name)
    raw = fn.__func__ if isinstance(fn, (staticmethod,)) else fn
    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            e = ACC.setdefault(key or name, [0.0, 0])
            e[0] += time.perf_counter() - t
            e[1] += 1
    setattr(obj, name, staticmethod(w) if isinstance(obj.__dict__.get(name), staticmethod) else w)
if REPS > 1:
    for o, n in ((q.KZG, "commit"), (q.KZG, "open_multilinear"), (q.KZG, "open_multilinear_begin"), (q.KZG, "open_multilinear_finish"),
                 (q.SumcheckProof, "prove"), (q.ZeroCheckProof, "prove"), (hp, "logup_denominators"), (hp, "small_int_table"),
                 (q.Context, "alloc"), (q.Context, "upload"), (q.Context, "eq_table"), (q.DeviceBuffer, "free")):
        if hasattr(o, n):
            _wrap(o, n, f"{getattr(o, '__name__', o)}.{n}")
def timed_loop_reps(fn, steps, warmup):
    for _ in range(warmup): fn()
    out = []
    for _ in range(REPS):
        ctx.sync(); t=time.perf_counter(); l0=ctx.kernel_launches
        fn()
        ctx.sync(); out.append(round((time.perf_counter()-t)*1e3, 1))
        print(out[-1], "ms:", {k: (round(v[0] * 1e3), v[1]) for k, v in sorted(ACC.items(), key=lambda kv: -kv[1][0])})
        ACC.clear()
    print("per-proof ms:", out)
    return min(out), ctx.kernel_launches-l0
r = bench.bench_hyperplonk(ctx, q, K, g, mont(bench.TAU), timed_loop_reps if REPS > 1 else timed_loop)
print({k: r[k] for k in ("value", "rows_per_trace", "gpu_launches")})
