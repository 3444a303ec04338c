"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total ms, mean us, share."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[iu], 1e-6)
    name = r[ik].split("(")[0].replace("void ", "").replace("mn::", "")
    tot[name][0] += 1; tot[name][1] += v
allms = sum(v[1] for v in tot.values())
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print("%-44s %6d launches %9.2f ms  mean %8.1f us  %5.1f %%" % (name[:44], n, ms, 1e3 * ms / n, 100 * ms / allms))
print("total %.2f ms" % allms)
