"""Per-kernel times of the last classifier training step in an ncu launch list (gpu__time_duration.sum csv of scripts/classifier_time.py)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
data = []
for r in rows[hdr + 1:]:
    try:
        data.append((r[ki].replace("hb::<unnamed>::", "").replace("hb::", "")[:60], float(r[vi].replace(",", "")), r[gi], r[bi]))
    except Exception:
        pass
idx = [i for i, d in enumerate(data) if d[0].startswith("adam_kernel")]
start, end = idx[-2] + 1, idx[-1]
tot = 0.0
for d in data[start:end + 1]:
    print(f"{d[1] / 1e3:8.1f} us  grid {d[2]:>14} block {d[3]:>12}  {d[0]}")
    tot += d[1]
print(f"{tot / 1e3:8.1f} us  sum of {end + 1 - start} launches (ncu: serialised, cold caches)")
