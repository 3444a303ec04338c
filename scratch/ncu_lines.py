"""Per-source-line totals (stall samples, instructions executed) from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = ""
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0] != "":      # a source line header row: totals of the line
        try:
            key = (fname, int(r[0]))
        except ValueError:
            continue
        d = agg.setdefault(key, [r[1].strip(), 0, 0])
        d[1] += int(r[4]) if r[4].isdigit() else 0; d[2] += int(r[7]) if r[7].isdigit() else 0
tot_s = sum(v[1] for v in agg.values()); tot_i = sum(v[2] for v in agg.values())
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
for (f, ln), (src, s_, i_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * s_ / tot_s, 100.0 * i_ / tot_i, f, ln, src[:110]))
