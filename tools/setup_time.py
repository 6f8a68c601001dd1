import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quill_zkvm_b200 as q
FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
mont = lambda v, mod=FR: np.frombuffer(((v % mod) * (1 << 256) % mod).to_bytes(32, "little"), dtype=np.uint8).copy()
ctx = q.Context(0)
g = np.concatenate([mont(1, FQ), mont(2, FQ)])
for logn in (20, 24):
    t = time.perf_counter(); kzg = q.KZG.trusted_setup(ctx, (1 << logn) - 1, g, mont(0x1234567)); ctx.sync(); t1 = time.perf_counter()
    kzg.precompute(); ctx.sync(); t2 = time.perf_counter()
    print(f"2^{logn}: srs_generate {1e3*(t1-t):.1f} ms, srs_precompute {1e3*(t2-t1):.1f} ms")
    kzg.srs.free()
