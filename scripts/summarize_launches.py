"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of each kernel."""
import collections, csv, io, sys
txt = open(sys.argv[1]).read()
r = csv.DictReader(io.StringIO(txt[txt.index('"ID"'):]))
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    name = row["Kernel Name"]
    if "spin_kernel" in name:  # bench.py holds the stream back with torch.cuda._sleep while it enqueues the event-timed steps
        continue
    name = name[:name.index("(")] if "(" in name else name
    agg[name[:90]][0] += 1
    agg[name[:90]][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'share':>6} {'n':>4} {'avg us':>9}  kernel   (total {tot:.0f} us over {sum(v[0] for v in agg.values())} launches)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / tot * 100:5.1f}% {v[0]:4d} {v[1] / v[0]:9.1f}  {k}")
