"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name: count, total us, share."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
tot = collections.defaultdict(float); cnt = collections.Counter()
order = []
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1000 if unit in ("ns", "nsecond") else v * (1 if unit in ("us", "usecond") else 1000)
    tot[name] += us; cnt[name] += 1
    order.append((name, us, row.get("Grid Size", ""), row.get("Block Size", "")))
T = sum(tot.values())
print(f"total {T:.1f} us over {sum(cnt.values())} launches")
for n, t in sorted(tot.items(), key=lambda kv: -kv[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{t:10.1f} us {100*t/T:5.1f}%  x{cnt[n]:4d}  {n[:110]}")
if len(sys.argv) > 3:
    for i, o in enumerate(order):
        print(i, o)
