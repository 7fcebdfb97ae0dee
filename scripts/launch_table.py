"""Condenses an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` log into one
line per launch.  python scripts/launch_table.py gpurun_out/launches.csv [max_rows]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
out = OrderedDict()
for r in rows[1:]:
    d = dict(zip(h, r))
    k = d["ID"]
    e = out.setdefault(k, {"name": d["Kernel Name"][:48], "grid": d["Grid Size"]})
    e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for k, e in list(out.items())[:lim]:
    print(f"{e['name']:50s} {e['grid']:>14s} {e.get('gpu__time_duration.sum', 0) / 1e3:9.1f} us  "
          f"R {e.get('dram__bytes_read.sum', 0) / 1e6:8.1f} MB  W {e.get('dram__bytes_write.sum', 0) / 1e6:8.1f} MB")
