"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel:
share of the step, launches, mean/min duration.  python tools/summarize_launches.py in.csv > out.md"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    try:
        agg[row["Kernel Name"]].append(float(row["Metric Value"].replace(",", "")))
    except (ValueError, KeyError):
        continue
tot = sum(sum(v) for v in agg.values())
print(f"total kernel time in capture: {tot/1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches "
      "(ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes)\n")
print("| share | launches | mean us | min us | kernel |")
print("|---:|---:|---:|---:|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| {sum(v)/tot*100:.2f}% | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {min(v)/1e3:.2f} | `{k}` |")
