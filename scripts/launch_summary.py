"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of device time per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:80]
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"total {tot/1e3:.2f} ms over {sum(n for n,_ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t/tot*100:6.2f}% {t/1e3:9.3f} ms  n={n:4d} avg={t/n:9.1f} us  {k}")
