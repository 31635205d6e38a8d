"""Per-kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: launch_summary.py list.csv"""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    n = r[ki].split("(")[0]
    c, t = tot.get(n, (0, 0.0)); tot[n] = (c + 1, t + v)
total = sum(t for _, t in tot.values())
print(f"{'kernel':<74}{'count':>6}{'total us':>11}{'share':>8}")
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{n[:72]:<74}{c:>6}{t:>11.1f}{100 * t / total:>7.1f}%")
