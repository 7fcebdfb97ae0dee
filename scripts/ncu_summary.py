"""Condense an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) into the handful of
counters the roofline discussion uses."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__cluster_size', 'smsp__inst_executed.sum',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_wait']
print("| kernel | grid | " + " | ".join(w.replace("smsp__pcsamp_warps_issue_stalled_", "stall:") for w in want if w in idx) + " |")
print("|---|---|" + "---|" * len([w for w in want if w in idx]))
for r in data:
    vals = []
    for w in want:
        if w in idx:
            v = r[idx[w]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {units[idx[w]]}".strip())
    print(f"| {r[idx['Kernel Name']][:48]} | {r[idx['launch__grid_size']]} | " + " | ".join(vals) + " |")
