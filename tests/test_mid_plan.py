"""Host logic of the persistent short-round kernel (csrc/sumcheck.cu mid_make_plan), checked without a GPU through the
test hook qz_test_mid_plan: for every round of an sc_mid launch the plan says how many blocks work on it, how many
pairs each takes, and which blocks may leave -- the kernel's grid-wide waits are only deadlock-free if these agree."""
import ctypes as C

import pytest

from quill_zkvm_b200 import _lib

ROUNDS = 40
TAIL_LOG = 11  # SC_TAIL_LOG: a sharded proof gathers its shards when G * size <= 2^11


def plan(size, pending, k, d, cap, G):
    lib = _lib.load()
    nblk, fut, chunk = ((C.c_uint32 * ROUNDS)() for _ in range(3))
    tile = C.c_uint32()
    grid = lib.qz_test_mid_plan(size, pending, k, d, cap, G, nblk, fut, chunk, C.byref(tile))
    return grid, list(nblk), list(fut), list(chunk), tile.value


def rounds_of(size, pending, G):
    """(pairs, after_gather) per round, replaying sc_mid's loop"""
    out, gathered = [], G == 1
    while True:
        if not gathered and size * G <= (1 << TAIL_LOG):
            size *= G
            gathered = True
        if size <= 1 or (pending and size == 2):
            return out
        out.append((size // 4 if pending else size // 2, gathered))
        if pending:
            size //= 2
        pending = 1


@pytest.mark.parametrize("G", [1, 2, 8])
@pytest.mark.parametrize("k,d", [(1, 1), (3, 3), (4, 4), (8, 7), (9, 2), (2, 9)])
@pytest.mark.parametrize("cap", [1, 148, 256])
def test_plan_covers_every_pair_and_never_strands_a_block(G, k, d, cap):
    for log_size in (1, 2, 5, 11, 12, 16, 18):
        for pending in (0, 1):
            size = 1 << log_size
            if size < G and G > 1:
                continue
            grid, nblk, fut, chunk, tile = plan(size, pending, k, d, cap, G)
            rs = rounds_of(size, pending, G)
            assert 1 <= grid <= cap and grid == max([1] + nblk[: len(rs)])
            fits = k <= 8 and d + 1 <= 8
            assert (tile > 0) == fits and (tile == 0 or 2 * k * tile <= 512)
            for j, (pairs, _) in enumerate(rs):
                assert 1 <= nblk[j] <= grid
                assert fut[j] == max(nblk[j:len(rs)]), "a block may only leave when no later round needs it"
                if chunk[j]:  # split round: contiguous chunks, every block non-empty, all pairs covered
                    assert fits and chunk[j] <= 128
                    assert (nblk[j] - 1) * chunk[j] < pairs <= nblk[j] * chunk[j]
                    if pairs <= 128:
                        assert nblk[j] == 1  # two tiles on one block beat gathering several blocks' vectors
                else:  # whole pairs per thread, strided over nblk * 256 threads
                    assert nblk[j] <= max(1, (pairs + 255) // 256)
            for j in range(len(rs), ROUNDS):
                assert nblk[j] == 1 and fut[j] == 1


def test_plan_rejects_bad_input():
    lib = _lib.load()
    a = (C.c_uint32 * ROUNDS)()
    t = C.c_uint32()
    assert lib.qz_test_mid_plan(16, 0, 3, 3, 0, 1, a, a, a, C.byref(t)) < 0
    assert lib.qz_test_mid_plan(16, 0, 3, 3, 8, 0, a, a, a, C.byref(t)) < 0
