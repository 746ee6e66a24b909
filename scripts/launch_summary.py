"""Summarise an `ncu --metrics gpu__time_duration.sum,... --csv` launch list per kernel (profiles/*_launches_*.csv)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[idi], {"name": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for v in per.values():
    n = v["name"].split("(")[0][:64]
    a = agg[n]
    a[0] += 1
    a[1] += v.get("gpu__time_duration.sum", 0)
    a[2] += v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0)
    a[3] += v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * v.get("gpu__time_duration.sum", 0)
tot = sum(a[1] for a in agg.values())
print(f"launches {len(per)}, total {tot / 1e6:.2f} ms")
print("| kernel | launches | time ms | share | DRAM GB | tensor pipe active (time-weighted) |")
print("|---|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"| `{n}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f} % | {a[2] / 1e9:.2f} | {a[3] / max(a[1], 1):.1f} % |")
