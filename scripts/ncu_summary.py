"""Summarise ncu captures into the small tracked files under profiles/.

  python scripts/ncu_summary.py launches <launch-list.csv> <out.json>     per-kernel count / total / share from a
                                                                          `--metrics gpu__time_duration.sum` launch list
  python scripts/ncu_summary.py full <raw.csv> <out.csv>                   selected columns of `ncu -i rep --page raw --csv`
"""
import collections
import csv
import json
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        us = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        name = r[ki].split("(")[0].replace("void ", "").replace("dfb::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    out = {"source": src, "total_us": total,
           "kernels": [{"kernel": k, "launches": a[0], "total_us": round(a[1], 2), "avg_us": round(a[1] / a[0], 2),
                        "share": round(a[1] / total, 4)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump(out, open(dst, "w"), indent=1)


def full(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    w = csv.writer(open(dst, "w"))
    w.writerow([hdr[c] for c in cols])
    w.writerow([units[c] for c in cols])
    for r in data:
        w.writerow([r[c] for c in cols])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
