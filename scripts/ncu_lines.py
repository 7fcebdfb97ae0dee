"""Per-source-line hot spots of one kernel from an ncu report: joins the SASS page of `ncu --page source --csv` with the
line table of the cubin (nvdisasm --print-line-info), instruction by instruction.
usage: python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX CUBIN MANGLED_SUBSTRING [top] [instance]
(KERNEL_REGEX matches ncu's base kernel name; `instance` picks among several captured launches / instantiations)"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, kre, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
instance = int(sys.argv[6]) if len(sys.argv) > 6 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for k, i in enumerate(starts):
    print(f"instance {k}: {rows[i - 1][1][:90]}", file=sys.stderr)
start = starts[instance]
end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
hdr = rows[start]
data = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
lines, cur, infn, fn_lines = [], None, False, []
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = mangled in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        fn_lines.append(cur)
if len(fn_lines) != len(data):
    print(f"warning: {len(fn_lines)} instructions in the cubin vs {len(data)} in the report", file=sys.stderr)
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r, where in zip(data, fn_lines):
    a = agg[where]
    a[0] += int(float(r[col["Instructions Executed"]] or 0))
    a[1] += int(float(r[col["# Samples"]] or 0))
    for s in stall_cols:
        v = int(float(r[col[s]] or 0))
        if v:
            a[2][s] += v
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
for where, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ", ".join(f"{k[6:]} {v}" for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:4])
    print(f"{str(where):34s} inst {100 * a[0] / tot_i:5.1f}%  samples {100 * a[1] / tot_s:5.1f}%  [{st}]")
