"""Condense an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / bench.py cite."""
import csv, io, subprocess, sys, json
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
lines = ["# ncu summary of `%s`" % rep.split("/")[-1], "",
         "Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200; read with",
         "`ncu -i ... --page raw --csv`.  One row per profiled launch.", ""]
lines.append("| kernel | " + " | ".join(w.replace(".avg.pct_of_peak_sustained", " %").replace("_elapsed", "").replace("_active", "") for w in want) + " | top stalls (per issue) |")
lines.append("|---|" + "---|" * (len(want) + 1))
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    vals = []
    for w in want:
        v = r[idx[w]] if w in idx else ""
        try:
            v = "%.4g" % float(v.replace(",", ""))
        except Exception:
            pass
        vals.append(v + (" " + units[idx[w]] if w in idx and units[idx[w]] not in ("", "%") and "bytes" in w or w.startswith("gpu__time") else ""))
    st = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            x = float(r[idx[h]])
            if x > 0.3:
                st.append((x, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    st = ", ".join("%s %.2f" % (n, x) for x, n in sorted(st, reverse=True)[:4])
    lines.append("| `%s` | " % name + " | ".join(vals) + " | " + st + " |")
open(out, "w").write("\n".join(lines) + "\n")
print(open(out).read()[:1500])
