"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr, agg = None, collections.defaultdict(list)
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            agg[d["Kernel Name"][:70]].append(float(d["Metric Value"].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) for v in agg.values()) or 1.0
print(f"{'kernel':70s} {'n':>5s} {'avg_us':>10s} {'total_ms':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} {len(v):5d} {sum(v)/len(v)/1e3:10.2f} {sum(v)/1e6:10.3f} {sum(v)/tot:7.3f}")
