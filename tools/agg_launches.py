"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= vi: continue
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    if r[ui] == "us": v *= 1e3
    elif r[ui] == "ms": v *= 1e6
    name = r[ki].split("(")[0][-70:]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} n={v[0]:6d} total={v[1]/1e6:9.3f} ms avg={v[1]/v[0]/1e3:8.2f} us share={v[1]/tot:.3f}")
