"""Per-kernel SASS opcode histogram of libmnemcu.so (cuobjdump -sass), the mnemonics that show what the hardware runs:
DMMA (FP64 tensor MMA), UBLKCP (TMA 1-D bulk copy), SYNCS (mbarrier), R2P + predicated DADD (masked FP64 add), DADD/DFMA/
DMUL/DSETP (FP64 pipe), LDS/STS (shared memory), SHFL, LDG/STG, ATOM/RED.  usage: sass_hist.py <lib.so> <out.md>"""
import collections, re, subprocess, sys
lib, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(2).split(".")[0]
        hist[kern][op] += 1
        if m.group(1) and op == "DADD":
            hist[kern]["@P DADD"] += 1
cols = ["DMMA", "UBLKCP", "SYNCS", "R2P", "@P DADD", "DADD", "DFMA", "DMUL", "DSETP", "LDS", "STS", "SHFL", "LDG", "STG", "ATOM", "ATOMS", "RED", "BAR"]
with open(out, "w") as f:
    f.write("# SASS opcode histogram of `%s` (sm_100a cubin, `cuobjdump -sass`)\n\n" % lib.split("/")[-1])
    f.write("Static instruction counts per kernel.  DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64), UBLKCP = TMA 1-D bulk copy\n"
            "(cp.async.bulk), SYNCS = mbarrier operations, R2P + `@P DADD` = the branch-free masked FP64 add of the E-step and the\n"
            "dictionary tables.  FP64 has no tcgen05 path, so no UTCMMA / LDTM is expected.\n\n")
    f.write("| kernel | total | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
    for k, h in hist.items():
        f.write("| `%s` | %d | " % (k[:70], sum(v for q, v in h.items() if q != "@P DADD")) + " | ".join(str(h.get(c, 0)) for c in cols) + " |\n")
print(open(out).read()[:3000])
