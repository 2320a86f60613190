"""Summarise an ncu --csv launch list (gpu__time_duration.sum): last K launches, name + us."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = [r for r in rows[h + 1:] if len(r) > vi and r[0].isdigit()]
k = int(sys.argv[2]) if len(sys.argv) > 2 else len(data)
tot = 0.0
for r in data[-k:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui].startswith("n") else v
    tot += v
    print(f"{v:9.1f} us  {r[ki][:100]}")
print(f"{tot:9.1f} us total over {min(k, len(data))} launches")
