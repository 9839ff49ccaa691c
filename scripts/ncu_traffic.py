"""profiles/rNN_ncu_traffic.json from an `ncu --set full --page raw --csv` export: per kernel the DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches, and the last launch -- for the Krylov kernels the
one with the most basis columns) next to duration and DRAM throughput.  usage: ncu_traffic.py raw.csv out.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}


def val(r, name):
    return float(r[col[name]].replace(",", ""))


agg = collections.OrderedDict()
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("dfb::", "")
    rd = val(r, "dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]]
    wr = val(r, "dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]]
    us = val(r, "gpu__time_duration.sum") * tscale[units[col["gpu__time_duration.sum"]]]
    thr = val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    agg.setdefault(name, []).append((rd + wr, us, thr))
out = {"source": sys.argv[1]}
for k, v in agg.items():
    out[k] = {"launches": len(v), "dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "us_per_launch": sum(x[1] for x in v) / len(v),
              "dram_throughput_pct": sum(x[2] for x in v) / len(v),
              "last_launch": {"dram_bytes": v[-1][0], "us": v[-1][1], "dram_throughput_pct": v[-1][2]}}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: (round(v["dram_bytes_per_launch"] / 1e6, 1), round(v["us_per_launch"], 1)) for k, v in out.items() if k != "source"}))
