"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X ...`):
    python tools/summarize_launches.py gpurun_out/X.csv "<comment>" > profiles/rNN_launch_summary.csv"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    ms = v / 1e6 if r[ui].startswith("n") else (v / 1e3 if r[ui].startswith("u") else v)
    name = re.sub(r"\(.*$", "", r[ki]).replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    tot += ms
print(f"# ncu launch list summary: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# cold-cache, serialised, un-throttled per-launch times: compare SHARES with the live run, not absolutes")
print(f"# total {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches")
print("kernel,launches,total_ms,share,avg_us")
for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"\"{k}\",{n},{ms:.2f},{ms / tot:.4f},{ms / n * 1e3:.1f}")
