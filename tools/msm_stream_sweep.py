"""Times KZG commit at 2^log_n (precomputed windows) for several segment settings of the streamed MSM:
QZ_MSM_SEGMENTS (host scalars, end to end) and QZ_MSM_SEGMENTS_DEV (device-resident scalars)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quill_zkvm_b200 as q  # noqa: E402

FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
TAU = 0x1234567890ABCDEF1234567890ABCDEF


def mont(v, mod=FR):
    return np.frombuffer(((v % mod) * (1 << 256) % mod).to_bytes(32, "little"), dtype=np.uint8).copy()


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << log_n
    stream = torch.cuda.Stream()
    ctx = q.Context(0, stream.cuda_stream)
    kzg = q.KZG.trusted_setup(ctx, n - 1, np.concatenate([mont(1, FQ), mont(2, FQ)]), mont(TAU)).precompute()
    dev = ctx.random_fr(n, 1)
    pin = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
    host = pin.numpy()
    host[:] = dev.download()
    host = host.reshape(-1, 32)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps=4, warm=2):
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
    for var, arg, settings in (("QZ_MSM_SEGMENTS_DEV", dev, ["1", "1,1"]),
                               ("QZ_MSM_SEGMENTS", host, ["1", "1,3,9", "1,4,12", "1,2", "1,3", "1,7", "1,2,5", "1,5"])):
        for s in settings:
            os.environ[var] = s
            ms, acc, r = timed(lambda: kzg.commit(arg))
            if ref is None:
                ref = r
            assert np.array_equal(r, ref), (var, s)
            print(f"{var}={s:12s} {ms:8.3f} ms/commit   accumulate {acc:7.3f} ms   launches/commit n/a", flush=True)
        os.environ.pop(var)
    ctx.close()


if __name__ == "__main__":
    main()
