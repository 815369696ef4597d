"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and time share per kernel."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= vi:
        continue
    try:
        ns = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0][:60]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ns
unit = rows[rows.index(hdr) + 1][hdr.index("Metric Unit")] if "Metric Unit" in hdr else "ns"
tot = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values())} launches, total {tot:.0f} {unit}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / tot:7.2%}  {n:5d} launches  {t / n:12.0f} {unit} each  {k}")
