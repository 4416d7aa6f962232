"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and average
duration and share per kernel (peak-measurement kernels listed but excluded from the share).
    python tools/launch_summary.py profiles/launches_r2.csv "<command that was profiled>" > profiles/launches_r2_summary.md"""
import csv, sys, collections
path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = r[4]
    tot[name] += float(r[14]); cnt[name] += 1
prod = sum(v for k, v in tot.items() if "_peak_kernel" not in k)
print("# ncu --metrics gpu__time_duration.sum --clock-control none, command: %s" % cmd)
print("# %d launches captured; per-launch times are cold-cache and serialised: compare SHARES, not absolutes." % len(rows))
print("kernel | launches | total ns | share of product kernels | avg ns")
for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    share = "-" if "_peak_kernel" in name else "%.4f" % (v / prod)
    print("%s | %d | %d | %s | %d" % (name, cnt[name], v, share, v / cnt[name]))
