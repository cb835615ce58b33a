"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python scripts/summarize_launches.py profiles/xxx_launches.csv > profiles/xxx_launches_summary.txt"""
import collections, csv, re, sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
h = rows[0]; idx = {n: i for i, n in enumerate(h)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) < len(h) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
    v = float(r[idx["Metric Value"]].replace(",", "")); unit = r[idx["Metric Unit"]]
    us = v / 1000 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1000)
    tot[name] += us; cnt[name] += 1
T = sum(tot.values())
print(f"launches {sum(cnt.values())}  total {T / 1000:.3f} ms (serialised, cold-cache per-launch times)")
print(f"{'kernel':50s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, v in tot.most_common():
    print(f"{k[:50]:50s} {cnt[k]:8d} {v / 1000:10.3f} {100 * v / T:6.1f}% {v / cnt[k]:9.1f}")
