"""Turn gpurun_out/ ncu artefacts into small committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1_d.csv > profiles/r1_launches.md
    python profiles/summarize.py full gpurun_out/prof_r1_d.ncu-rep   > profiles/r1_ncu_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    order = []
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        t = t / 1e3 if unit == "ns" else (t * 1e3 if unit == "ms" else t)
        order.append((re.sub(r"\(.*", "", row["Kernel Name"])[:70], t))
    n = len(order)
    # one step = from a patchify launch (first kernel of the ViT) to the next one; take the last complete step
    starts = [i for i, (s, _) in enumerate(order) if "patchify" in s]
    q = order[starts[-2]:starts[-1]] if len(starts) >= 2 else order[-(n // 4):]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, t in q:
        agg[s][0] += 1
        agg[s][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"ncu launch list `{path}`: {n} launches captured; table = the last step ({len(q)} launches, "
          f"{tot / 1e3:.2f} ms of serialised, cold-cache kernel time — compare SHARES, not absolutes)\n")
    print("| share | time (us) | launches | avg (us) | kernel |\n|---|---|---|---|---|")
    for s, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {100 * t / tot:.1f}% | {t:.1f} | {c} | {t / c:.1f} | `{s}` |")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__inst_executed.sum"]
    print(f"`ncu --set full --clock-control none` capture `{path}` (raw page, selected metrics)\n")
    print("| # | kernel | " + " | ".join(w.split(".")[0].replace("__", " ") for w in want) + " |")
    print("|---|---|" + "---|" * len(want))
    for d in data:
        cells = [f"{d[idx[w]]} {units[idx[w]]}" if w in idx else "-" for w in want]
        print(f"| {d[idx['ID']]} | `{d[idx['Kernel Name']][:48]}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
