"""profiles/search_ncu.json (read by bench.py's `roofline_search`) from an .ncu-rep of search_persistent_kernel."""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
r = next(x for x in rows[2:] if "search_persistent_kernel" in x[idx["Kernel Name"]])
f = lambda k: float(r[idx[k]].replace(",", ""))
d = {"source": rep.split("/")[-1], "kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", ""),
     "ms_under_ncu": f("gpu__time_duration.sum") / (1e6 if "ns" in rows[1][idx["gpu__time_duration.sum"]] else 1.0) if rows[1][idx["gpu__time_duration.sum"]] in ("ns", "nsecond") else f("gpu__time_duration.sum"),
     "fp64_pipe_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
     "issue_slots_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
     "warp_instructions": f("smsp__inst_executed.sum"),
     "registers_per_thread": f("launch__registers_per_thread"), "grid": f("launch__grid_size"), "block": f("launch__block_size"),
     "dram_bytes": f("dram__bytes_read.sum") + f("dram__bytes_write.sum"), "note": note}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d))
