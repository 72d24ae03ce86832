"""python tools/ncu_summary.py launches.csv: per-kernel launch count / average / share from an
`ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[1:]:
    us = float(r[vi].replace(",", "")) / {"ns": 1000.0, "us": 1.0, "ms": 1e-3, "s": 1e-6}.get(r[ui], 1000.0)
    d.setdefault(r[ki][:70], []).append(us)
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print(f"{k:72s} n={len(v):4d} avg={sum(v) / len(v):10.1f} us  share={100 * sum(v) / tot:5.1f} %")
