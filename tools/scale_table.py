"""Markdown table of the strong-scaling runs: python tools/scale_table.py N=file [N=file ...]  (bench.py JSON lines)"""
import json
import sys

rows = []
for arg in sys.argv[1:]:
    n, path = arg.split("=")
    d = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    rows.append((int(n), d))
rows.sort()
base = rows[0][1]


def eff(t1, tn, n):
    return f"{t1 / tn:.1f}×" if n > 1 else ""


print("| GPUs | MSM (device scalars) | MSM (host scalars, e2e) | sumcheck 3 × 2²⁴ | zero-check 3 × 2²⁴ | HyperPlonk 2 × 2²⁰ rows |")
print("|---|---|---|---|---|---|")
for n, d in rows:
    hp = d.get("hyperplonk_prove", {}).get("value")
    hp1 = base.get("hyperplonk_prove", {}).get("value")
    zc, zc1 = d.get("zerocheck", {}).get("ms_per_step"), base.get("zerocheck", {}).get("ms_per_step")
    print(f"| {n} | {d['ms_per_step']:.2f} ms {eff(base['ms_per_step'], d['ms_per_step'], n)} | {d['e2e']['ms_per_step']:.2f} ms | "
          f"{d['sumcheck']['ms_per_step']:.2f} ms {eff(base['sumcheck']['ms_per_step'], d['sumcheck']['ms_per_step'], n)} | "
          f"{zc:.2f} ms {eff(zc1, zc, n)} | " + (f"{hp:.2f} s {eff(hp1, hp, n)} |" if hp else "— |"))
same = all(r[1].get("digests") == base.get("digests") for r in rows)
print("\ndigests identical across N:", same)
