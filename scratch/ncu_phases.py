"""Instruction / stall-sample shares per phase of search.cu from the cuda,sass source CSV (line ranges given as name:lo-hi)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ranges = [(a.split(":")[0], int(a.split(":")[1].split("-")[0]), int(a.split(":")[1].split("-")[1])) for a in sys.argv[2:]]
fname = ""; tot = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] == "" or not r[0].isdigit(): continue
    ln = int(r[0]); smp = int(r[4]) if r[4].isdigit() else 0; ins = int(r[7]) if r[7].isdigit() else 0
    key = "other:" + fname
    if fname == "search.cu":
        key = "search.cu:other"
        for name, lo, hi in ranges:
            if lo <= ln <= hi: key = name; break
    t = tot.setdefault(key, [0, 0]); t[0] += smp; t[1] += ins
S = sum(v[0] for v in tot.values()); I = sum(v[1] for v in tot.values())
for k, (s_, i_) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %5.1f%% inst %5.1f%% samples" % (k, 100.0 * i_ / I, 100.0 * s_ / S))
