"""Writes profiles/roofline_traffic.json from an `ncu --set full` capture of one bench step.

    ncu --set full --clock-control none -o gpurun_out/step python scripts/one_step.py          (on the GPU box)
    ncu -i gpurun_out/step.ncu-rep --page raw --csv > profiles/r2_step_raw.csv                 (here)
    python scripts/ncu_traffic.py profiles/r2_step_raw.csv --config 2 --regime clustered --launches-per-step 8

bench.py reports `roofline.traffic` and `stage.dram_bytes_per_step` from this file as long as its source hash matches
the CUDA sources the library is built from (bench.source_hash); a stale file yields null, never an old number."""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("raw_csv")
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--regime", default="clustered")
ap.add_argument("--launches-per-step", type=int, default=8)
ap.add_argument("--roofline-kernel", default="roialign_fwd_kernel<2, 1>")
ap.add_argument("--note", default="")
a = ap.parse_args()

rows = list(csv.reader(open(a.raw_csv)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
units = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]


def to_bytes(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


step = data[-a.launches_per_step:]
per_kernel = []
total = 0.0
roof = None
for r in step:
    b = to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    name = r[col["Kernel Name"]]
    per_kernel.append({"kernel": name[:80], "dram_bytes": int(b),
                       "duration_us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) *
                       {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)})
    total += b
    if a.roofline_kernel.replace(" ", "") in name.replace(" ", ""):
        roof = int(b)
path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
try:
    rec = json.load(open(path))
except Exception:
    rec = {}
src = bench.source_hash()
if rec.get("source_sha256") != src:
    rec = {"source_sha256": src, "configs": {}}
rec["configs"][f"{a.config}:{a.regime}"] = {"roofline_kernel_dram_bytes": roof, "step_dram_bytes": int(total),
                                            "kernels": per_kernel, "from": os.path.basename(a.raw_csv), "note": a.note}
json.dump(rec, open(path, "w"), indent=1)
print(f"{path}: config {a.config}:{a.regime} roofline kernel {roof} B, step {int(total)} B over {len(step)} launches")
