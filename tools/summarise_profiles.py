"""Turn the ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/."""
import collections
import csv
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
REPS = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/"   # where the .ncu-rep files are
OUT = sys.argv[3] if len(sys.argv) > 3 else "profiles/"      # where the text summaries go


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


def hot_spots(rep, out):
    """append, per captured launch, where the warps' stall samples fall by opcode (ncu source page, SASS view)"""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    kern, hdr, data = None, None, collections.OrderedDict()
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            kern = f"{len(data)}: {r[1][:90]}"
            data[kern] = []
            hdr = None
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and kern and len(r) == len(hdr):
            data[kern].append(dict(zip(hdr, r)))
    o, seen = [], set()
    for k, v in data.items():
        tot = sum(int(x.get("# Samples") or 0) for x in v) or 1
        sig = (k.split(": ", 1)[1], len(v), tot)  # the page lists a launch once per view: keep the first
        if sig in seen:
            continue
        seen.add(sig)
        by, cnt = collections.Counter(), collections.Counter()
        for x in v:
            parts = x["Source"].split()
            if not parts:
                continue
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            key = op.split(".")[0] + (".WIDE" if "WIDE" in op else "")
            by[key] += int(x.get("# Samples") or 0)
            cnt[key] += 1
        o.append(f"## stall samples by opcode, launch {k}  ({len(v)} SASS instructions, {tot} samples)")
        for op, smp in by.most_common(8):
            o.append(f"   {op:12s} {cnt[op]:6d} instructions  {100 * smp / tot:5.1f} % of the samples")
    open(out, "a").write("\n".join(o) + "\n")


if __name__ == "__main__":
    import os
    import shutil
    G = "gpurun_out/"
    os.makedirs(OUT, exist_ok=True)
    shutil.copyfile(G + "launches.csv", f"{OUT}{TAG}_launches_bench_steps2.csv")
    launches(G + "launches.csv", f"{OUT}{TAG}_launches_summary.txt",
             f"{TAG}: ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 2 --warmup 1 "
             "--no-cpu-baseline --mlpcs-log-n 0 --hyperplonk-log-rows 0   (MSMs of 2^24 with device-resident and with host scalars -- streamed in 3 ranges --, precomputed windows; sumcheck and zero-check proofs of 3 x 2^24; setup)")
    if os.path.exists(G + "lhp20.csv"):
        launches(G + "lhp20.csv", f"{OUT}{TAG}_launches_hyperplonk_2_20.txt",
                 f"{TAG}: ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 python tools/profile_hp.py 20   "
                 "(setup + 2 HyperPlonk proofs of two 2^20-row traces: BASELINE config 5 on one GPU)")
    CAP = [("prof_msm", "msm_accumulate", "-k regex:msm_accumulate -c 1 python tools/profile_one.py msm 24 pre  (2^24 points, c = 22, 12 mixed additions per point)"),
           ("prof_bucket_reduce", "msm_bucket_reduce", "-k regex:msm_bucket_reduce -c 1 python tools/profile_one.py msm 24 pre  (2^21 buckets, 64 per thread)"),
           ("prof_sort", "msm_sort_onesweep", "-k regex:Onesweep -c 3 python tools/profile_one.py msm 24 pre  (cub::DeviceRadixSort over 12 x 2^24 (key, value) pairs, 22-bit keys: three onesweep passes)"),
           ("prof_sc", "sc_round_prod", "-k regex:sc_round_prod -c 3 python tools/profile_one.py sumcheck 24  (launch 0 = round 0, evaluate only; launches 1, 2 = rounds 1, 2, fold fused, X = 1 derived)"),
           ("prof_sc_mid", "sc_mid", "-k regex:sc_mid -c 1 python tools/profile_one.py sumcheck 24  (the 18 rounds from 2^18 entries down, one cooperative launch)"),
           ("prof_zc", "sc_round_zc", "-k regex:sc_round_zc -c 2 python tools/profile_one.py zerocheck 24  (eq-factored zero-check: round 0 and round 1)"),
           ("prof_ntt", "ntt_pass", "-k regex:ntt_pass -c 3 python tools/profile_one.py mlpcs 22  (forward 2^23 transform: passes of 8 + 8 + 7 stages)")]
    for rep, name, what in CAP:
        if os.path.exists(REPS + rep + ".ncu-rep"):
            full(REPS + rep + ".ncu-rep", f"{OUT}{TAG}_ncu_{name}.txt", f"{TAG}: ncu --set full --clock-control none {what}")
            hot_spots(REPS + rep + ".ncu-rep", f"{OUT}{TAG}_ncu_{name}.txt")
