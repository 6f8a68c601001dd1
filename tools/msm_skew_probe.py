"""Commit times for scalar distributions a prover meets: random, small integers (id / permutation tables), 0/1 selectors."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quill_zkvm_b200 as q
from quill_zkvm_b200.hyperplonk import small_int_table
FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
def mont(v, mod=FR):
    return np.frombuffer(((v % mod) * (1 << 256) % mod).to_bytes(32, "little"), dtype=np.uint8).copy()
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << log_n
stream = torch.cuda.Stream()
ctx = q.Context(0, stream.cuda_stream)
kzg = q.KZG.trusted_setup(ctx, n - 1, np.concatenate([mont(1, FQ), mont(2, FQ)]), mont(0x1234567)).precompute()
cases = {"random": ctx.random_fr(n, 1),
         "ids 1..n": ctx.upload(small_int_table(ctx, np.arange(1, n + 1, dtype=np.uint64))),
         "selector 0/1": ctx.upload(small_int_table(ctx, (np.arange(n, dtype=np.uint64) % 3 == 0).astype(np.uint64))),
         "first row only": ctx.upload(small_int_table(ctx, (np.arange(n, dtype=np.uint64) == 0).astype(np.uint64)))}
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, buf in cases.items():
    with torch.cuda.stream(stream):
        kzg.commit(buf); torch.cuda.synchronize(); ev0.record(stream)
        for _ in range(3): kzg.commit(buf)
        ev1.record(stream); torch.cuda.synchronize()
    print(f"{name:16s} {ev0.elapsed_time(ev1) / 3:8.3f} ms/commit", flush=True)
ctx.close()
