"""Turn the ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/."""
import collections
import csv
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def launches(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        name = row["Kernel Name"].split("(")[0][-56:]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    o = [f"# {title}", "# cold-cache, serialised per-launch times: compare SHARES, not absolutes", "total_ms  launches  share  kernel"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        o.append(f"{v:10.3f} {cnt[k]:6d} {100 * v / T:6.2f}%  {k}")
    open(out, "w").write("\n".join(o) + "\n")


KEEP = ["Kernel Name", "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "inst_executed",
        "sm__inst_executed.avg.per_cycle_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    o = [f"# {title}"]
    for n, row in enumerate(rows[2:]):
        o.append(f"## captured launch {n}")
        for h, u, v in zip(hdr, units, row):
            if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                o.append(f"{h} [{u}] = {v}")
    open(out, "w").write("\n".join(o) + "\n")


if __name__ == "__main__":
    import shutil
    shutil.copyfile("gpurun_out/launches.csv", f"profiles/{TAG}_launches_bench_steps2.csv")
    launches("gpurun_out/launches.csv", f"profiles/{TAG}_launches_summary.txt",
             f"{TAG}: ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 2 --warmup 1 "
             "--no-cpu-baseline --mlpcs-log-n 0 --hyperplonk-log-rows 0   (3 MSMs of 2^24 with device-resident scalars and 3 with host scalars -- streamed in 3 ranges --, precomputed windows; 6 sumcheck proofs of 3 x 2^24; setup)")
    full("gpurun_out/prof_r1_msm.ncu-rep", f"profiles/{TAG}_ncu_msm_accumulate.txt",
         f"{TAG}: ncu --set full --clock-control none -k regex:msm_accumulate -c 1 python tools/profile_one.py msm 24 pre  (2^24 points, c = 22, 12 mixed additions per point)")
    import os
    if os.path.exists("gpurun_out/lhp20.csv"):
        launches("gpurun_out/lhp20.csv", f"profiles/{TAG}_launches_hyperplonk_2_20.txt",
                 f"{TAG}: ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 python tools/profile_hp.py 20   "
                 "(setup + 2 HyperPlonk proofs of two 2^20-row traces: BASELINE config 5 on one GPU)")
    if os.path.exists("gpurun_out/prof_r1_ntt.ncu-rep"):
        full("gpurun_out/prof_r1_ntt.ncu-rep", f"profiles/{TAG}_ncu_ntt_pass.txt",
             f"{TAG}: ncu --set full --clock-control none -k regex:ntt_pass -c 3 python tools/profile_one.py mlpcs 22  "
             "(forward 2^23 transform: passes of 8 + 8 + 7 stages; the first two gather one twiddle per butterfly)")
    full("gpurun_out/prof_r1_sc.ncu-rep", f"profiles/{TAG}_ncu_sc_round_prod.txt",
         f"{TAG}: ncu --set full --clock-control none -k regex:sc_round_prod -c 2 python tools/profile_one.py sumcheck 24  (launch 0 = round 0, evaluate only; launch 1 = round 1, fold fused)")
