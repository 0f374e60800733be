"""Turns gpurun_out/{launches.csv,*.ncu-rep} into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py <tag> [launches.csv] [report.ncu-rep]
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "profiles"
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active",
        "gpu__dram_throughput", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__warps_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem",
        "lts__throughput", "lts__t_bytes", "lts__t_sector_hit_rate", "l1tex__throughput", "sm__throughput",
        "sm__cycles_elapsed", "smsp__cycles_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "smsp__warp_issue_stalled", "launch__occupancy", "sm__ctas_launched")


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else (v * 1e3 if u in ("s", "second") else v))
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += ms
        tot += ms
    out = [f"# ncu launch list — {tag}", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).", "",
           "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k[:110]}` | {n} | {ms:.3f} | {100 * ms / tot:.2f} % |")
    out.append(f"| **total** | | {tot:.3f} | |")
    (OUT / f"{tag}_launches.md").write_text("\n".join(out) + "\n")
    print("\n".join(out[:14]))


def full(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name") or any(h.startswith(k) for k in KEEP)]
    with open(OUT / f"{tag}_full.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{j}" for j in range(len(rows) - 2)])
        for i in cols:
            w.writerow([hdr[i], units[i]] + [r[i] for r in rows[2:]])
    print("wrote", OUT / f"{tag}_full.csv", len(cols), "metrics x", len(rows) - 2, "launches")


if __name__ == "__main__":
    tag = sys.argv[1]
    OUT.mkdir(exist_ok=True)
    if len(sys.argv) > 2 and Path(sys.argv[2]).exists():
        launches(tag, sys.argv[2])
    if len(sys.argv) > 3 and Path(sys.argv[3]).exists():
        full(tag, sys.argv[3])
