#!/usr/bin/env python
"""bench.py -- the EM hot path of M&NEM on B200: EM iterations/s, candidate-graph scores/s, E-step HBM GB/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 5] [--starts S] [--impl ours|reference]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one complete EM phase of mnem(D, phi = init): every EM start of the config (sharded over the ranks,
N > 1: the ranks drain one shared queue of starts, mnem_b200.dist.StartQueue; D replicated) runs do_inits to convergence (R/mnems.r:1884-2165).  `value` = EM loop
iterations summed over all starts and ranks / device time of the step (CUDA events inside the library, max over
ranks), with D already resident in HBM.  `e2e` times the same step through the C ABI from HOST buffers: the H2D copy
of D (pinned memory), the layout pass, the EM phase and the D2H copy of the results are all inside the timed region.
The default workload is BASELINE.json's scaling config (30 S-genes x 2000 E-genes x 1M cells, k = 5, 100 starts).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# EM loop iterations per start of the seeded benchmark fits.  The CPU arm cannot afford to run 100 starts at 1M cells to
# convergence (about an hour of one core per start), so it evaluates its measured per-iteration cost at this many
# iterations per start.  The numbers are the means over all starts of the device fits of the same seeded inputs
# (profiles/r2_bench_cfg*.json, "em_iters_per_step" / starts); the device's iteration counts are those of the oracle
# on every start the oracle has replayed (tests/test_gpu_scale_parity.py: n_iter bit-exact).  bench.py's own
# `cpu_baseline` leg (N = 1) does not use the table: it takes the mean of the fit it has just timed.
MEAN_EM_ITERS = {1: 4.0, 2: 5.3, 3: 11.98, 4: 10.0, 5: 13.07}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--starts", type=int, default=0, help="override the config's number of EM starts")
    ap.add_argument("--cells", type=int, default=0, help="override the config's number of cells")
    ap.add_argument("--batch", type=int, default=0, help="EM starts per pass over D (0: k*batch <= 30)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-cells", type=int, default=0, help="cells in the CPU-baseline sample (0: auto)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def make_workload(args):
    from mnem_b200 import simdata
    cfg = dict(simdata.CONFIGS[args.config])
    if args.cells:
        cfg["C"] = args.cells
    if args.starts:
        cfg["starts"] = args.starts
    sim = simdata.make_config(cfg, seed=1000 + args.config)
    init = simdata.init_phi(sim["S"], sim["k"], sim["starts"], seed=2000 + args.config)
    name = "cfg%d: simData-style S=%d E=%d C=%d k=%d starts=%d complete=%s" % (
        args.config, sim["S"], sim["E"], sim["C"], sim["k"], sim["starts"], sim["complete"])
    return sim, init, name


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (R cannot be installed in this image)
# ---------------------------------------------------------------------------------------------------------------
def cpu_sample(sim, init, n_threads, cells, guard=1):
    """One bounded run of the CPU restatement: the first `cells` cells of the same data, one EM start per thread, INIT
    phase + exactly `guard` EM loop iterations each (max_iter_guard), with the reference's redundant recomputation
    switched on (faithful_cost).  Returns wall seconds and counts."""
    import oracle as orc
    cells = min(cells, sim["C"])
    D = orc.simgen(sim, 0, cells, nthreads=max(1, n_threads))
    pert = sim["pert"][:cells]
    ph = init[:n_threads]
    t0 = time.time()
    r = orc.mnem_em(D, pert, sim["S"], ph, complete=sim["complete"], nthreads=n_threads, faithful_cost=True,
                    max_iter_guard=guard)
    dt = time.time() - t0
    return dict(seconds=dt, em_iters=int(r["n_iter"].sum()), cand_scored=int(r["n_scored"].sum()), cells=cells,
                starts=len(ph))


def cpu_extrapolated(sim, init, n_threads, cells, mean_iters, c0=0, greedy_comps=None):
    """EM iterations/s of the CPU restatement for the FULL workload, from bounded component timings.

    Per start the reference spends (R/mnems.r:1905-1915, 1976-2108; the oracle's orc_em_start with faithful_cost):
        T(C) = getProbs(theta = NULL)(C)  +  n * [ 2k * doMean(C)  +  2k * nem_greedy  +  getProbs(theta given)(C) ]
    The three data passes are timed on the first `cells` cells of the same matrix and scale linearly with the number of
    cells (they are plain loops over the columns); the greedy searches work on the collapsed E x S matrix and do not
    depend on C.  Every thread times its own start (the threads run concurrently, so memory-bandwidth contention is in
    the numbers); the value reported is  n_threads * n / T(C_full)  with n = `mean_iters` EM iterations per start."""
    import concurrent.futures

    import oracle as orc
    c = min(cells, sim["C"])
    c0 = max(0, min(c0, sim["C"] - c))
    k, S, E = sim["k"], sim["S"], sim["E"]
    D = orc.simgen(sim, c0, c, nthreads=max(1, n_threads))       # the oracle's own generator twin: no product library here
    pert = sim["pert"][c0:c0 + c]
    n_threads = max(1, min(n_threads, len(init)))

    def one(s):
        phi = init[s]
        t = {}
        t0 = time.time()
        g0 = orc.get_probs(D, pert, S, phi, None, np.zeros((k, c)), None, complete=sim["complete"], faithful_cost=True)
        t["init"] = time.time() - t0
        post = orc.get_affinity(g0["probs"], g0["mw"], complete=sim["complete"])
        t0 = time.time()
        R = [orc.do_mean(D, pert, S, post[i]) for i in range(k)]
        t["domean"] = (time.time() - t0) / k
        # both restarts (empty start, previous graph) of the first `ng` components are timed; the searches do not depend on
        # the number of cells, so a short step may time fewer of the 2k searches of an EM iteration (their mean is used)
        ng = k if greedy_comps is None else max(1, min(k, greedy_comps))
        t0 = time.time()
        fits = [orc.nem_greedy(R[i], None) for i in range(ng)] + [orc.nem_greedy(R[i], phi[i]) for i in range(ng)]
        t["greedy"] = (time.time() - t0) / (2 * ng)
        nsc = sum(f["n_scored"] for f in fits)
        adj = np.stack([fits[i]["adj"] if i < ng else phi[i] for i in range(k)])
        th = np.stack([fits[i]["theta"] if i < ng else g0["theta_used"][i] for i in range(k)])
        t0 = time.time()
        orc.get_probs(D, pert, S, adj, th, g0["probs"], g0["mw"], complete=sim["complete"], faithful_cost=True)
        t["estep"] = time.time() - t0
        t["scored"] = nsc
        t["n_searches"] = 2 * ng
        return t

    t0 = time.time()
    with concurrent.futures.ThreadPoolExecutor(max_workers=n_threads) as ex:      # ctypes releases the GIL
        ts = list(ex.map(one, range(n_threads)))
    wall = time.time() - t0
    m = {key: float(np.mean([t[key] for t in ts])) for key in ("init", "domean", "greedy", "estep")}
    scale = sim["C"] / float(c)
    t_init = m["init"] * scale
    t_em = 2 * k * m["domean"] * scale + 2 * k * m["greedy"] + m["estep"] * scale
    T = t_init + mean_iters * t_em
    value = n_threads * mean_iters / T
    return dict(value=float(value), seconds=wall, t_init_full=t_init, t_em_full=t_em, cells=c, starts=n_threads,
                mean_iters=mean_iters,
                cand_scores_per_s=float(sum(t["scored"] for t in ts) / max(1e-9, sum(t["n_searches"] * t["greedy"] for t in ts)) * n_threads),
                seconds_greedy=float(np.mean([t["n_searches"] * t["greedy"] for t in ts])),
                sample=("oracle port (C restatement of cbg-ethz/mnem 1.23.0, not R), %d start(s) concurrently, cells %d..%d of %d "
                        "(%.1f s): per start getProbs(theta=NULL) %.2f s, doMean %.3f s, getProbs(theta) %.2f s (each scaled by "
                        "%d/%d cells), nem_greedy %.2f s per search (independent of the cell count); redundant passes kept; "
                        "T = init + n*(2k*doMean + 2k*greedy + getProbs) at n = %.2f EM iterations per start: init %.3g s, "
                        "EM iteration %.3g s per start" % (n_threads, c0, c0 + c, sim["C"], wall, m["init"], m["domean"], m["estep"],
                                                          sim["C"], c, m["greedy"], mean_iters, t_init, t_em)))


def bench_config(args, sim, name):
    """The `config` object of the JSON line: the same for both arms."""
    return {"workload": name, "batch": args.batch or max(1, min(30 // sim["k"], sim["starts"])),
            "l2": "inputs larger than L2 (D = %.1f GB per layout)" % (8e-9 * sim["E"] * sim["C"]),
            "candidate_count": "graphs generated and scored, before unique()"}


def run_reference(args):
    """The reference's CPU implementation of the path (the oracle port: R is not installable here) on all host threads.

    One step = one bounded sample of the workload: one EM start per host thread, all concurrently, on a slice of C/32
    cells -- a DIFFERENT slice and different starts in every step, so the W + K steps of a default run together time
    6/32 of the cells of the matrix once each instead of one small sample over and over.  Each step times the four
    component passes of an EM iteration (cpu_extrapolated) and reports EM iterations/s for the full workload; what is
    extrapolated (linear in the cell count for the three data passes) is spelled out in `sample`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sim, init, name = make_workload(args)
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so OpenMP's own count is not asked;
    # the samples run in a Python thread pool, ctypes releases the GIL)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    nthr = max(1, min(ncpu, sim["starts"]))
    # Size of a step.  The whole run (W + K steps) should end within a few minutes whatever K is (the driver asks for 25
    # steps): MNEM_REF_BUDGET_S seconds (default 300) are spread over the steps.  The first step takes a small slice and
    # measures the cost per cell; the following steps take slices sized to the per-step budget, at most C / (W + K) cells so
    # that the slices stay disjoint, at most C / 32.
    n_steps = max(1, args.warmup + args.steps)
    per_step = float(os.environ.get("MNEM_REF_BUDGET_S", "300")) / n_steps
    cap = max(2000, min(sim["C"] // 32, sim["C"] // n_steps))
    cells = args.cpu_cells or max(2000, min(cap, sim["C"] // 256))
    few = per_step < 25.0                         # short steps time 2 of the 2k searches of an EM iteration
    vals, c0 = [], 0
    for i in range(n_steps):
        first = (i * nthr) % max(1, len(init) - nthr + 1)
        t0 = time.time()
        r = cpu_extrapolated(sim, init[first:first + nthr], nthr, cells, MEAN_EM_ITERS.get(args.config, 10.0), c0=c0,
                             greedy_comps=1 if few else None)
        took = time.time() - t0
        c0 += cells
        r["cells_step"], r["seconds_step"] = cells, took
        if i >= args.warmup:
            vals.append(r)
        if not args.cpu_cells:
            per_cell = max(1e-9, (took - r["seconds_greedy"]) / cells)
            cells = int(max(2000, min(cap, (per_step - r["seconds_greedy"]) / per_cell)))
    v = float(np.mean([x["value"] for x in vals]))
    sample = "%d timed steps on disjoint slices, the last: %s" % (len(vals), vals[-1]["sample"])
    line = {"impl": "reference", "metric": "em_iterations_per_s", "value": v, "unit": "EM iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean([x["seconds"] for x in vals])), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, sim, name),
            "values_per_step": [x["value"] for x in vals], "cells_per_step": [x["cells_step"] for x in vals],
            "seconds_per_step_wall": [x["seconds_step"] for x in vals],
            "cpu_baseline": {"value": v, "unit": "EM iterations/s", "cores": nthr, "kind": "port", "sample": sample},
            "cand_scores_per_s": float(np.mean([x["cand_scores_per_s"] for x in vals])),
            "e2e": {"value": v, "unit": "EM iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import mnem_b200 as mb
    from mnem_b200 import dist as mdist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sim, init, name = make_workload(args)
    k, E, Cn, S = sim["k"], sim["E"], sim["C"], sim["S"]
    launches0 = mb.launch_count()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(xs):
        t = torch.tensor(xs, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    # ---- value: D resident in HBM ----
    cx = mb.Context.synthetic(sim, device=local)
    recs = []
    sampler = ClockSampler(local)
    launches_timed = 0
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            sampler.start()
            launches_timed = mb.launch_count()
        queue = mdist.StartQueue(sim["starts"])          # N = 1: a local counter
        barrier()
        t0 = time.time()
        r = mb.em_run(cx, init, complete=sim["complete"], batch=args.batch, want_probs=False, next_start=queue)
        torch.cuda.synchronize()
        wall = time.time() - t0
        barrier()
        my = sorted(queue.taken)
        st = r["stats"]
        ms = reduce_max(st["ms_total"])
        tot = reduce_sum([st["em_iters"], st["cand_scored"], st["greedy_iters"]])
        if i >= args.warmup:
            recs.append(dict(ms=ms, wall=reduce_max(wall), em_iters=tot[0], cand=tot[1], st=st))
        winner, lls = mdist.gather_winner(my, r["ll"][my], sim["starts"], device=dev)
    clocks = sampler.stop()
    launches_timed = mb.launch_count() - launches_timed
    ms_step = float(np.mean([x["ms"] for x in recs]))
    em_iters = float(np.mean([x["em_iters"] for x in recs]))
    value = em_iters / (ms_step * 1e-3)
    st = recs[-1]["st"]
    peak, peak_kind = peaks()
    t_estep = st["ms_estep"] / max(1, st["passes_estep"]) * 1e-3
    bytes_launch = st["estep_bytes"] / max(1, st["passes_estep"])
    achieved = bytes_launch / t_estep / 1e9 if t_estep > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "estep_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "estep_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "traffic": traffic, "algorithmic_bytes_per_launch": bytes_launch, "ms_per_launch": t_estep * 1e3,
                "launches_per_step": st["passes_estep"],
                "pairs_per_launch": (bytes_launch - 8.0 * E * Cn) / (8.0 * Cn),
                "share_of_step": st["ms_estep"] / max(st["ms_total"], 1e-9)}
    # above 20 pairs per launch the pass is bound by the FP64 add rate before HBM (one predicated DADD per pair, E-gene and
    # cell; 64 FP64 lanes per SM and clock): report that roofline next to the HBM one
    pairs = roofline["pairs_per_launch"]
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    mhz = clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965.0
    fp64_peak = sms * 64.0 * mhz * 1e6
    fp64_rate = pairs * E * float(Cn) / t_estep if t_estep > 0 else 0.0
    roofline["fp64_adds"] = {"achieved": fp64_rate, "peak": fp64_peak, "unit": "FP64 adds/s", "frac": fp64_rate / fp64_peak,
                             "peak_source": "%d SMs x 64 lanes x %.0f MHz (sampled SM clock)" % (sms, mhz)}
    roofline["note"] = ("%.0f pairs per pass: above 20 pairs the pass is bound by the FP64 add rate (fp64_adds.frac), not by HBM; "
                        "with 20 pairs per pass (--batch 4 at k = 5) the same kernel runs at 0.89 of the HBM peak but the whole fit is ~8 %% "
                        "slower (153 against 166 EM iterations/s when both were measured, profiles/README.md)" % pairs)
    # the kernel that decides `value`: the device-resident greedy search (one persistent launch per EM tick).  Live numbers
    # from this run; pipe utilisation from the committed ncu capture of the same kernel (profiles/search_ncu.json).
    cand_rank0 = st["cand_scored"] / max(1e-9, st["ms_search"] * 1e-3)
    roofline_search = {"bound": "fp64 issue (shared-memory look-ups + FP64 compares)", "kernel": "search_persistent_kernel",
                       "share_of_step": st["ms_search"] / max(st["ms_total"], 1e-9),
                       "candidates_per_s_in_search_rank0": cand_rank0, "sm_busy_in_kernel": st.get("search_sm_busy", 0.0),
                       "ms_search_per_step": st["ms_search"]}
    sp = os.path.join(ROOT, "profiles", "search_ncu.json")
    if os.path.exists(sp):
        roofline_search["ncu"] = json.load(open(sp))
    breakdown = {"ms_estep": st["ms_estep"], "ms_mstats": st["ms_mstats"], "ms_search": st["ms_search"],
                 "ms_mixture": st["ms_mixture"], "ms_total": st["ms_total"], "passes_estep": st["passes_estep"],
                 "passes_mstats": st["passes_mstats"], "search_sm_busy": st.get("search_sm_busy", 0.0)}
    cand_rate = float(np.mean([x["cand"] for x in recs])) / (ms_step * 1e-3)
    cand_rate_search = float(np.sum(reduce_sum([st["cand_scored"]]))) / max(1e-9, reduce_max(st["ms_search"]) * 1e-3)
    cx.close()

    # ---- e2e: host buffers in, host buffers out ----
    e2e = None
    if not args.no_e2e:
        # every rank holds its own pinned host copy of D (the caller's matrix); ranks agree on go / no-go first so that a
        # rank that cannot allocate does not leave the others waiting at a barrier
        host, err = None, None
        try:
            import psutil
            need = 8.0 * E * Cn
            if psutil.virtual_memory().available < (1.5 + world) * need:
                raise MemoryError("not enough host memory for a pinned copy of D per rank")
            host = torch.empty((Cn, E), dtype=torch.float64, pin_memory=True)       # column-major E x C
        except Exception as ex:                                                    # noqa: BLE001
            err = ex
        if err is None:
            free_dev = torch.cuda.mem_get_info(dev)[0]
            want_dev = 8.0 * E * Cn * (3.2 if world > 1 else 2.2)       # two layouts (+ the gathered copy for N > 1)
            if free_dev < want_dev:
                err = MemoryError("not enough device memory for the end-to-end run: %.1f GB free" % (free_dev / 1e9))
        if -reduce_max(-(1.0 if err is None else 0.0)) < 1.0:
            e2e = {"value": None, "unit": "EM iterations/s", "error": repr(err)[:200] if err else "another rank failed"}
        else:
            hnp = host.numpy()
            Dh = hnp.T
            with mb.Context.synthetic(sim, device=local) as c0:
                mb._lib.check(mb.load_library().mnemcu_ctx_read_data(c0._h, 0, Cn, mb._lib.dp(hnp)))
            times, iters, parts = [], [], []
            sampler2 = ClockSampler(local)
            sampler2.start()
            # N > 1: every rank uploads 1/N of the columns from ITS host copy and the slices are exchanged over NVLink
            # (one NCCL all-gather) instead of pushing N full copies of D through PCIe and the host memory bus
            per = -(-Cn // world)
            for i in range(max(1, min(args.warmup, 1)) + int(os.environ.get("MNEM_E2E_STEPS", max(1, min(args.steps, 2))))):
                queue = mdist.StartQueue(sim["starts"])
                barrier()
                t0 = time.time()
                if world == 1:
                    cxe = mb.Context(Dh, sim["pert"], device=local)
                else:
                    # one in-place all-gather over NVLink: rank r owns rows [r*per, (r+1)*per) of the (padded) device matrix
                    Dd = torch.empty((world * per, E), dtype=torch.float64, device=dev)
                    mine = Dd[rank * per:(rank + 1) * per]
                    n_mine = max(0, min(Cn, (rank + 1) * per) - rank * per)
                    mine[:n_mine].copy_(host[rank * per:rank * per + n_mine], non_blocking=True)
                    dist.all_gather_into_tensor(Dd, mine)
                    torch.cuda.synchronize()
                    cxe = mb.Context.from_device(Dd.data_ptr(), E, Cn, sim["pert"], device=local)
                    del Dd
                torch.cuda.synchronize()
                t_up = time.time() - t0
                with cxe:
                    # probs="winner": the k x C log ratios of the best start come back to the host, as in mnem()'s result
                    r = mb.em_run(cxe, init, complete=sim["complete"], batch=args.batch, probs="winner", next_start=queue)
                    t_em = time.time() - t0 - t_up
                my = sorted(queue.taken)
                mdist.gather_winner(my, r["ll"][my], sim["starts"], device=dev)     # the winner is part of the result
                torch.cuda.synchronize()
                dt = reduce_max(time.time() - t0)
                barrier()
                if i >= 1:
                    times.append(dt)
                    parts.append((t_up, t_em, r["stats"]["ms_total"] * 1e-3))
                    iters.append(reduce_sum([r["stats"]["em_iters"]])[0])
            per_start = k * S * S + 4 * k * E + 8 * k + 8 + 4 + 4 * 8 * 256          # adj, theta, mw, ll, n_iter, traces
            e2e = {"value": float(np.mean(iters) / np.mean(times)), "unit": "EM iterations/s",
                   "h2d_bytes_per_step": int(8 * E * Cn + world * (4 * Cn + sim["starts"] * k * S * S)),
                   "nvlink_bytes_per_step": int(8 * E * Cn * (world - 1)),
                   "d2h_bytes_per_step": int(sim["starts"] * per_start + world * 8 * k * Cn),   # + every rank's best probs
                   "seconds_per_step": float(np.mean(times)), "steps": len(times), "seconds_each": [float(x) for x in times],
                   "clocks": sampler2.stop(),
                   "seconds_upload_and_layout_rank0": float(np.mean([p_[0] for p_ in parts])),
                   "seconds_upload_each_rank0": [float(p_[0]) for p_ in parts],
                   "seconds_em_wall_rank0": float(np.mean([p_[1] for p_ in parts])),
                   "seconds_em_device_rank0": float(np.mean([p_[2] for p_ in parts]))}
            del hnp, Dh
        del host

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cells = args.cpu_cells or max(2000, Cn // 128)
        c = cpu_extrapolated(sim, init, 1, cells, em_iters / max(1, sim["starts"]))
        cpu = {"value": c["value"], "unit": "EM iterations/s", "cores": 1, "kind": "port", "sample": c["sample"],
               "cand_scores_per_s": c["cand_scores_per_s"]}
    if rank == 0:
        line = {"metric": "em_iterations_per_s", "value": value, "unit": "EM iterations/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": bench_config(args, sim, name), "starts_on_rank0": len(my),
                "sharding": "one queue of starts drained by all ranks" if world > 1 else "single GPU",
                "em_iters_per_step": em_iters, "cand_scores_per_s": cand_rate,
                "cand_scores_per_s_in_search": cand_rate_search, "estep_gbs": achieved,
                "roofline": roofline, "roofline_search": roofline_search, "breakdown_last_step_rank0": breakdown, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches_timed), "clocks": clocks, "winner_start": int(winner),
                "wall_ms_per_step": 1e3 * float(np.mean([x["wall"] for x in recs]))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
