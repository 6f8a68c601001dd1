"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Sharded MSM and sharded sumcheck (product fast path, generic expression, several sizes around the gather threshold)
must equal the oracle's unsharded results byte for byte on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quill_zkvm_b200 as q  # noqa: E402
from oracle import coracle as co  # noqa: E402
from oracle import pyref as py  # noqa: E402
from quill_zkvm_b200 import parallel  # noqa: E402
from tests import util  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = q.Context(local)
    parallel.init_comm(ctx)
    FR = py.FR
    tau, gen = 0xABCDEF12345, py.g1_mul(py.G1_GEN, 11)
    # ---- MSM ----
    for n in (world, 1000, 1 << 14):
        srs = co.srs_generate(co.g1_to_bytes(gen), co.fr1(tau), n, threads=4)
        sc = util.rand_fr(n, 5 + n)
        lo, hi = parallel.shard_range(n, rank, world)
        kz = q.KZG.from_points(ctx, srs[lo:hi]) if hi > lo else q.KZG.from_points(ctx, np.zeros((0, 64), np.uint8))
        got = kz.msm_sharded(sc[lo:hi])
        want = co.msm(srs, sc, mode=1, threads=4)
        assert np.array_equal(got, want), f"rank {rank}: sharded MSM n={n} differs"
        kz.srs.free()
        # the same product with the whole SRS and scalar vector on every rank (qz_msm_split), host and device scalars
        kz = q.KZG.from_points(ctx, srs)
        assert np.array_equal(kz.commit_split(sc), want), f"rank {rank}: split MSM n={n} differs"
        d = ctx.upload(sc)
        assert np.array_equal(kz.commit_split(d), want), f"rank {rank}: split MSM (device scalars) n={n} differs"
        d.free()
        kz.srs.free()
    ctx.comm_resync()  # collective re-agreement on the mailbox sequence numbers: the exchanges below must still line up
    # ---- sumcheck ----
    exprs = [util.expr_product(3), util.expr_from_py(py.e_sub(py.e_mul(py.e_in(0), py.e_in(1)), py.e_mul(py.e_const(9), py.e_in(2))))]
    for nv in (3, 8, 11, 12, 13, 16, 18, 19, 20, 22):  # 19, 20: sc_mid's multi-block rounds exchange with the peers; 22: streaming rounds first
        if (1 << nv) < world:
            continue
        tabs = [util.rand_fr(1 << nv, 77 * nv + t) for t in range(3)]
        lo, hi = parallel.table_shard_range(nv, rank, world)
        for nodes, consts in exprs:
            store = q.VirtualPolynomialStore(nv)
            store.polynomials = [np.ascontiguousarray(t[lo:hi]) for t in tabs]
            h = store.new_virtual_from_expr(util.to_qexpr(nodes, consts))
            tr = q.Transcript(b"multi", ctx)
            sc, claim = q.SumcheckProof.prove(ctx, nv, store, h, co.fr1(3), tr, sharded=True)
            st = co.transcript_new(b"multi")
            o = co.sumcheck_prove(nv, tabs, nodes, consts, co.fr1(3), st, max_coeffs=8, threads=4)
            for j in range(nv):
                assert np.array_equal(sc.r_polys[j], o["coeffs"][j][: o["lens"][j]]), f"rank {rank} nv {nv} round {j}"
            assert np.array_equal(claim.point, o["point"]) and np.array_equal(claim.evaluation, o["evaluation"])
            assert tr.state.tobytes() == st.tobytes()
    # ---- zero-check: eq-factored rounds (product) and the streamed eq shard (generic expression) ----
    for nv in (8, 12, 13, 16, 20):
        if (1 << nv) < world:
            continue
        tabs = [util.rand_fr(1 << nv, 55 * nv + t) for t in range(3)]
        lo, hi = parallel.table_shard_range(nv, rank, world)
        for nodes, consts in exprs:
            store = q.VirtualPolynomialStore(nv)
            store.polynomials = [np.ascontiguousarray(t[lo:hi]) for t in tabs]
            h = store.new_virtual_from_expr(util.to_qexpr(nodes, consts))
            tr = q.Transcript(b"multi_zc", ctx)
            proof, claim = q.ZeroCheckProof.prove(ctx, store, h, tr, sharded=True)
            st = co.transcript_new(b"multi_zc")
            o = co.sumcheck_prove(nv, tabs, nodes, consts, None, st, max_coeffs=q._lib.QZ_MAX_ROUND_COEFFS, zerocheck=True, threads=4)
            for j in range(nv):
                assert np.array_equal(proof.sumcheck_proof.r_polys[j], o["coeffs"][j][: o["lens"][j]]), f"rank {rank} zc nv {nv} round {j}"
            assert np.array_equal(proof.z, o["z"]) and np.array_equal(claim.point, o["point"])
            assert np.array_equal(claim.evaluation, o["evaluation"]) and tr.state.tobytes() == st.tobytes()
    # ---- HyperPlonk: openings dealt to the ranks (hyperplonk.OpeningBatch), every rank ends with the whole proof ----
    from oracle import fastkzg  # noqa: E402
    from quill_zkvm_b200 import hyperplonk as hp  # noqa: E402
    from tests.test_gpu_hyperplonk import to_product_circuit  # noqa: E402

    c1, w1 = py.fibonacci_circuit_and_trace(64)
    c2, w2 = py.modified_fibonacci_circuit_and_trace()
    circuits, witnesses = [c1, c2], [w1, w2]
    max_degree = max(c.num_cols() * c.num_rows for c in circuits)
    fastkzg.install_fast_s_polynomial()
    want = py.hyperplonk_prove(circuits, witnesses, fastkzg.FastKZG(max_degree, gen, tau))
    kzg = q.KZG.trusted_setup(ctx, max_degree, co.g1_to_bytes(gen), co.fr1(tau))
    prover = hp.HyperPlonk.preprocess(ctx, [to_product_circuit(c) for c in circuits], kzg)
    got = util.hyperplonk_py(prover.prove(kzg, [[co.to_mont(col) for col in w] for w in witnesses]))
    kzg.srs.free()
    assert got["state_end"] == want["state_end"], f"rank {rank}: HyperPlonk transcript differs"
    assert got["witness_commitment"] == want["witness_commitment"]
    for g, w in zip(got["trace_proofs"], want["trace_proofs"]):
        for key in ("openings_zero_check", "openings_public", "opening_id", "opening_permutation", "opening_permutation_trace"):
            assert g[key] == w[key], f"rank {rank}: {key} differs"
        for key in ("opening_proof_denom_left", "opening_proof_denom_right", "r_polys"):
            assert g["permutation"][key] == w["permutation"][key], f"rank {rank}: permutation {key} differs"
    # ---- the same at a size where the sharded commits, the sharded zero-check / logup sumchecks and the pooled openings all
    # engage (2^14 rows; the Python oracle would need minutes): against the single-GPU prover on a context without a
    # communicator, which tests/test_gpu_hyperplonk.py holds to the oracle ----
    rows = 1 << 14
    In, C = q.VirtualPolyExpr.Input, hp.Const

    def fib_circuit():
        c = hp.TransitionCircuit(rows)
        s1, s2 = c.allocate_state_cell(), c.allocate_state_cell()
        c.enforce_boundary_constraint(0, In(s1[0]))
        c.enforce_boundary_constraint(0, hp.Sub(In(s2[0]), C(1)))
        c.enforce_constraint(hp.Sub(In(s2[1]), In(s1[0]) + In(s2[0])))
        c.enforce_constraint(hp.Sub(In(s1[1]), In(s2[0])))
        w = [[0] * rows for _ in range(c.num_cols())]
        a, b = 0, 1
        for r in range(rows):
            w[s1[0]][r], w[s2[0]][r] = a, b
            a, b = b, (a + b) % FR
            w[s1[1]][r], w[s2[1]][r] = a, b
        return c, [co.to_mont(col) for col in w]

    circ, wit = fib_circuit()
    proofs = []
    for c in (ctx, q.Context(local)):  # with the communicator, then alone
        kz = q.KZG.trusted_setup(c, circ.num_cols() * rows, co.g1_to_bytes(gen), co.fr1(tau)).precompute()
        pr = hp.HyperPlonk.preprocess(c, [circ], kz)
        proofs.append(util.hyperplonk_py(pr.prove(kz, [wit])))
        kz.srs.free()
        if c is not ctx:
            c.close()
    assert proofs[0] == proofs[1], f"rank {rank}: the 2^14-row HyperPlonk proof differs from the single-GPU prover's"
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU parity ok on {world} ranks: sharded MSM, sharded sumcheck, sharded zero-check and the HyperPlonk proof with its "
              f"openings dealt to the ranks match the oracle bit for bit "
              f"(exchange: {'peer mailboxes over NVLink' if ctx.peer_memory else 'NCCL all-gather'})")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
