"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST step."""
import csv, collections, re, sys
path = sys.argv[1]
last_n = int(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path, errors="ignore")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
data = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    data.append((name, v))
if last_n:
    data = data[-last_n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in data:
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(a[1] for a in agg.values())
print("launches %d  total %.0f us" % (len(data), tot))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print("%9.0f us %5.1f%% %4d  %s" % (a[1], 100 * a[1] / tot, a[0], k[:100]))
if "--detail" in sys.argv:
    pat = sys.argv[sys.argv.index("--detail") + 1]
    for n, v in data:
        if pat in n:
            print("%8.1f  %s" % (v, n[:80]))
