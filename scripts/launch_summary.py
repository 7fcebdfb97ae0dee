"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: last step's kernels + per-kernel totals."""
import collections, csv, sys
path = sys.argv[1]
tail = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
def us(r): return float(r["Metric Value"].replace(",", "")) / 1e3
print(f"{len(rows)} launches; last {tail}:")
tot = 0
for r in rows[-tail:]:
    print(f"  {r['Kernel Name'][:64]:64s} grid {r['Grid Size']:>16s} block {r['Block Size']:>14s} {us(r):9.1f} us")
    tot += us(r)
print(f"  sum of the last {tail}: {tot:.1f} us")
agg, cnt = collections.defaultdict(float), collections.Counter()
for r in rows:
    k = r["Kernel Name"][:64] + " " + r["Grid Size"]
    agg[k] += us(r); cnt[k] += 1
total = sum(agg.values())
print("per kernel (all launches):")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"  {k:84s} n={cnt[k]:3d} avg {v / cnt[k]:9.1f} us  share {100 * v / total:5.1f} %")
