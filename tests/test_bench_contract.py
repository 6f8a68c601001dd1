"""The bench contract, checked without a GPU through the reference arm (`bench.py --impl reference`): one JSON line with
the keys the driver reads, the oracle port timed on the host cores, and no kernel launches claimed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-log-n", "10", "--ref-sc-log-n", "10"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["sumcheck"]["unit"] == "field-elems/s" and line["sumcheck"]["value"] > 0
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1", "--ref-log-n", "10", "--ref-sc-log-n", "10"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_traffic_comes_from_the_committed_profiles_and_config_is_shared():
    """roofline.traffic is read from the newest ncu summary under profiles/ (not a constant in bench.py); both arms of
    the bench print the same `config` object; the digest file the GPU arm compares itself with is complete."""
    import argparse
    import json

    import bench

    msm, msm_src = bench.profile_traffic("msm")
    sc, sc_src = bench.profile_traffic("sumcheck")
    assert msm_src.startswith("profiles/") and sc_src.startswith("profiles/")
    assert 14.5e9 < msm < 40e9      # algorithmic 14.5 GB; the 64-byte gathers pull 128-byte lines
    assert 6.0e9 < sc < 7.5e9       # algorithmic 6.44 GB: no re-reads
    args = argparse.Namespace(log_n=24, no_precompute=False)
    assert bench.workload_config(args, 1) == bench.workload_config(args, 1, peer_memory=False)  # one GPU: no exchange
    assert bench.workload_config(args, 8)["exchange"] != bench.workload_config(args, 8, peer_memory=False)["exchange"]
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_digests_2_24.json")))
    assert g["log_n"] == 24 and set(g["digests"]) >= {"commitment_xy", "commitment_serialized", "sumcheck_claimed_sum",
                                                      "sumcheck_final_transcript_state", "sumcheck_evaluation",
                                                      "zerocheck_final_transcript_state"}
    assert bench.shard_seed(5, 0) == 5 and bench.shard_seed(5, 1) == (5 + 4 * bench.GOLDEN64) % (1 << 64)
