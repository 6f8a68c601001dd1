"""Pair levels of the MSM (QZ_MSM_PAIR_LEVELS, csrc/msm.cu msm_pair_*): parity against the oracle at small sizes for
several level counts (every case is reported, nothing stops at the first mismatch), then timings of KZG commit at
2^log_n with precomputed windows, device-resident and host scalars, for 0..6 levels."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quill_zkvm_b200 as q  # noqa: E402
from oracle import coracle as co  # noqa: E402  (checker of the parity part only)

FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
TAU = 0x1234567890ABCDEF1234567890ABCDEF


def mont(v, mod=FR):
    return np.frombuffer(((v % mod) * (1 << 256) % mod).to_bytes(32, "little"), dtype=np.uint8).copy()


def rand_fr(n, seed):
    rng = np.random.default_rng(seed)
    vals = [int.from_bytes(rng.bytes(32), "little") % FR for _ in range(n)]
    return co.to_mont(vals)


def parity(ctx):
    g = np.concatenate([mont(1, FQ), mont(2, FQ)])
    kzg = q.KZG.trusted_setup(ctx, (1 << 16) - 1, g, mont(TAU))
    bases_all = kzg.srs.download()
    bad = 0
    for pre in (None, 0, 9):
        k = kzg if pre is None else q.KZG.from_points(ctx, bases_all).precompute(pre)
        for n in (64, 1000, 4097, 1 << 16):
            cases = {"random": rand_fr(n, n), "small": co.to_mont([(i * 7) % 5 for i in range(n)]),
                     "equal": co.to_mont([FR - 3] * n)}
            for name, sc in cases.items():
                want = co.msm(bases_all[:n], sc, mode=1, threads=os.cpu_count() or 1)
                for levels in (0, 1, 2, 3, 5, 8):
                    os.environ["QZ_MSM_PAIR_LEVELS"] = str(levels)
                    got = k.commit(sc)
                    ok = np.array_equal(got, want)
                    bad += not ok
                    if not ok or levels == 8:
                        print(f"parity pre={pre} n={n} {name} levels={levels}: {'ok' if ok else 'MISMATCH'}", flush=True)
        if pre is not None:
            k.srs.free()
    kzg.srs.free()
    os.environ.pop("QZ_MSM_PAIR_LEVELS", None)
    print("parity mismatches:", bad, flush=True)
    return bad


def timings(ctx, stream, log_n):
    n = 1 << log_n
    g = np.concatenate([mont(1, FQ), mont(2, FQ)])
    kzg = q.KZG.trusted_setup(ctx, n - 1, g, mont(TAU)).precompute()
    dev = ctx.random_fr(n, 1)
    pin = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
    host = pin.numpy()
    host[:] = dev.download()
    host = host.reshape(-1, 32)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps=3, warm=2):
        with torch.cuda.stream(stream):
            for _ in range(warm):
                r = fn()
            torch.cuda.synchronize()
            ev0.record(stream)
            acc = []
            for _ in range(steps):
                r = fn()
                acc.append(ctx.last_elapsed_ms(1))
            ev1.record(stream)
            torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / steps, sum(acc) / len(acc), r

    ref = None
    level_list = [int(x) for a in sys.argv if a.startswith("--levels=") for x in a.split("=")[1].split(",")] or [0, 2, 3, 4, 5, 6]
    for levels in level_list:
        os.environ["QZ_MSM_PAIR_LEVELS"] = str(levels)
        try:
            ms, acc, r = timed(lambda: kzg.commit(dev))
        except Exception as e:  # noqa: BLE001
            print(f"2^{log_n} levels={levels}: FAILED {e}", flush=True)
            continue
        if ref is None:
            ref = r
        same = np.array_equal(r, ref)
        line = f"2^{log_n} device scalars levels={levels}: {ms:8.3f} ms/commit  pair levels + accumulate {acc:7.3f} ms  same={same}"
        if levels in (0, 4):
            msh, _, rh = timed(lambda: kzg.commit(host))
            line += f"   host scalars {msh:8.3f} ms same={np.array_equal(rh, ref)}"
        print(line, flush=True)
    os.environ.pop("QZ_MSM_PAIR_LEVELS", None)
    dev.free()
    kzg.srs.free()


def main():
    stream = torch.cuda.Stream()
    ctx = q.Context(0, stream.cuda_stream)
    t0 = time.time()
    if "--no-parity" not in sys.argv:
        parity(ctx)
        print(f"parity part {time.time() - t0:.1f} s", flush=True)
    for a in sys.argv[1:]:
        if a.isdigit():
            timings(ctx, stream, int(a))
    ctx.close()


if __name__ == "__main__":
    main()
