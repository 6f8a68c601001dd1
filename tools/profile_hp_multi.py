"""Where a multi-GPU HyperPlonk proof spends its wall time, per kind of library call, on every rank:
    python -m torch.distributed.run --nproc-per-node N tools/profile_hp_multi.py K     (two traces of 2^K rows)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
import quill_zkvm_b200 as q
from quill_zkvm_b200 import hyperplonk as hp
from quill_zkvm_b200 import parallel

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
ctx = q.Context(local, stream.cuda_stream)
if world > 1:
    parallel.init_comm(ctx)
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
mont = lambda v, mod=bench.FR: np.frombuffer(((v % mod) * (1 << 256) % mod).to_bytes(32, "little"), dtype=np.uint8).copy()  # noqa: E731
ACC = {}


def wrap(obj, name):
    fn = getattr(obj, name)

    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            e = ACC.setdefault(f"{getattr(obj, '__name__', obj)}.{name}", [0.0, 0])
            e[0] += time.perf_counter() - t
            e[1] += 1
    setattr(obj, name, staticmethod(w) if isinstance(obj.__dict__.get(name), staticmethod) else w)


for o, n in ((q.KZG, "commit"), (q.KZG, "commit_split"), (q.KZG, "open"), (q.KZG, "open_multilinear_begin"), (q.SumcheckProof, "prove"),
             (q.ZeroCheckProof, "prove"), (hp, "logup_denominators"), (q.Context, "alloc"), (q.Context, "upload"), (q.Context, "allgather"),
             (q.Context, "field_op"), (q.Transcript, "append_bytes"), (q.Transcript, "draw_field_element")):
    if hasattr(o, n):
        wrap(o, n)


def timed_loop(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    out = []
    for _ in range(2):
        ACC.clear()
        ctx.sync()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        fn()
        ctx.sync()
        out.append((time.perf_counter() - t) * 1e3)
        sys.stdout.write(f"rank {rank}: {out[-1]:.1f} ms: " + str({k: (round(v[0] * 1e3, 1), v[1]) for k, v in sorted(ACC.items(), key=lambda kv: -kv[1][0])}) + "\n")
        sys.stdout.flush()
        if world > 1:
            dist.barrier()
            time.sleep(0.05 * rank)
    return min(out), 0


r = bench.bench_hyperplonk(ctx, q, K, np.concatenate([mont(1, FQ), mont(2, FQ)]), mont(bench.TAU), timed_loop)
if rank == 0:
    print({k: r[k] for k in ("value", "rows_per_trace")})
ctx.close()
if world > 1:
    dist.destroy_process_group()
