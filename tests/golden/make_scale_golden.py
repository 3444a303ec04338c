#!/usr/bin/env python
"""Golden vectors for the EM fit at BASELINE scale (tests/test_gpu_scale_parity.py).

  python tests/golden/make_scale_golden.py [--case NAME ...] [--threads 8]

Runs the CPU oracle (oracle/mnem_oracle.c, long-double accumulators = R on x86-64) on the same synthetic inputs that
bench.py / the GPU tests build on the device (mnem_b200.simdata: seeded tables + counter-based noise, bit-identical on
host and device) and stores what a fit returns: graphs, theta, log-likelihood traces, mixture weights, iteration and
candidate counts, the oracle's decision margins, and a strided sample of the k x C log-ratio matrix.  One fit runs on
`--threads` OpenMP threads that split independent outputs only (orc_set_inner_threads), so the numbers are those of the
serial oracle.  The oracle needs the full E x C matrix in host memory (16 GB for cfg5) and minutes to hours of CPU
time, which is why the vectors are committed instead of being recomputed inside the GPU test.

When an EM-level `ll_new > ll_kept` decision (R/mnems.r:2104) is at rounding level (margin <= 1e-12) the fit is
repeated with the oracle's noise_mode 1 and 2 so that both legitimate outcomes of `limits$mw` are on file.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

# name -> (BASELINE config, cells override (0 = the config's), EM starts taken from init_phi(seed = 2000 + cfg))
CASES = {
    "cfg3_full": (3, 0, 3),            # 20 S-genes x 1000 E-genes x 100 000 cells, k = 5
    "cfg4_full": (4, 0, 3),            # 24 x 1000 x 200 016, k = 3 (the perturb-seq shape; its default 3 starts)
    "cfg5_c62500": (5, 62500, 2),      # 30 x 2000 x 62 520, k = 5
    "cfg5_full": (5, 0, 1),            # 30 x 2000 x 1 000 020, k = 5: the north-star size, one start
}
SAMPLE = 4096                          # cells of `probs` kept per start


def build_case(name):
    from mnem_b200 import simdata
    cfgno, cells, starts = CASES[name]
    cfg = dict(simdata.CONFIGS[cfgno])
    if cells:
        cfg["C"] = cells
    cfg["starts"] = starts
    sim = simdata.make_config(cfg, seed=1000 + cfgno)
    init = simdata.init_phi(sim["S"], sim["k"], starts, seed=2000 + cfgno)
    return sim, init


def sample_index(C):
    return np.unique(np.linspace(0, C - 1, min(SAMPLE, C)).astype(np.int64))


def stage_weights(C):
    """Deterministic pseudo-responsibilities in [0, 1) for the doMean stage check (same formula on both sides)."""
    c = np.arange(C, dtype=np.uint64)
    return ((c * np.uint64(2654435761)) % np.uint64(1 << 32)).astype(np.float64) / float(1 << 32)


def row_index(E):
    return np.arange(0, E, 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", action="append", default=[])
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--f64", action="store_true",
                    help="also fit with double accumulators (R without long double): how far the reference is from itself")
    a = ap.parse_args()
    import oracle as orc
    from mnem_b200 import simdata
    for name in (a.case or list(CASES)):
        sim, init = build_case(name)
        t0 = time.time()
        D = simdata.host_data_native(sim, nthreads=a.threads)
        t_gen = time.time() - t0
        orc.set_inner_threads(a.threads)
        out = {}
        log = dict(case=name, S=sim["S"], E=sim["E"], C=sim["C"], k=sim["k"], starts=int(init.shape[0]),
                   complete=bool(sim["complete"]), threads=a.threads, seconds_generate=t_gen)
        idx = sample_index(sim["C"])
        modes = [0]
        t0 = time.time()
        ref = orc.mnem_em(D, sim["pert"], sim["S"], init, complete=sim["complete"], nthreads=1, noise_mode=0)
        log["seconds_fit"] = time.time() - t0
        if ref["min_accept_margin"].min() <= 1e-12:
            modes = [0, 1, 2]
        refs = {0: ref}
        for m in modes[1:]:
            refs[m] = orc.mnem_em(D, sim["pert"], sim["S"], init, complete=sim["complete"], nthreads=1, noise_mode=m)
        for m, r in refs.items():
            p = "m%d_" % m
            out[p + "mw"] = r["mw"]
            out[p + "final_mw"] = np.stack([orc.final_mw(r["probs"][s], r["mw"][s], complete=sim["complete"])
                                            for s in range(init.shape[0])])
            if m == 0:
                continue
            out[p + "ll"] = r["ll"]
        T = max(int(ref["n_iter"].max()), 1)

        def pad(tr):
            o = np.full((len(tr), T), np.nan)
            for s, v in enumerate(tr):
                o[s, :len(v)] = v
            return o
        out.update(adj=ref["adj"].astype(np.uint8), theta=ref["theta"].astype(np.int16), ll=ref["ll"], best=np.int64(ref["best"]),
                   n_iter=ref["n_iter"], n_scored=ref["n_scored"], n_raw=ref["n_raw"], n_greedy_iter=ref["n_greedy_iter"],
                   min_margin=ref["min_margin"], min_accept_margin=ref["min_accept_margin"],
                   lls=pad(ref["lls"]), phievo=pad(ref["phievo"]), thetaevo=pad(ref["thetaevo"]), mwevo=pad(ref["mwevo"]),
                   probs_idx=idx, probs_sample=ref["probs"][:, :, idx],
                   probs_rowsum=ref["probs"].sum(axis=2), probs_abs_rowsum=np.abs(ref["probs"]).sum(axis=2),
                   modes=np.array(modes), init=init.astype(np.uint8))
        out.update(n_dec=ref["n_dec"], n_dec_tight=ref["n_dec_tight"])
        # ---- stage vectors: every data pass at this size on inputs both sides can rebuild bit for bit ----
        S, E, Cn, k = sim["S"], sim["E"], sim["C"], sim["k"]
        t0 = time.time()
        gp = orc.get_probs(D, sim["pert"], S, ref["adj"][0], ref["theta"][0], np.zeros((k, Cn)), None, complete=sim["complete"])
        out.update(gp_ll=gp["ll"], gp_mw=gp["mw"], gp_probs_sample=gp["probs"][:, idx], gp_probs_rowsum=gp["probs"].sum(axis=1))
        g0 = orc.get_probs(D, sim["pert"], S, init[0], None, np.zeros((k, Cn)), None, complete=sim["complete"])
        out.update(g0_ll=g0["ll"], g0_mw=g0["mw"], g0_probs_sample=g0["probs"][:, idx], g0_theta=g0["theta_used"].astype(np.int16))
        w = stage_weights(Cn)
        R = orc.do_mean(D, sim["pert"], S, w)
        out.update(dm_rows=row_index(E), dm_R=R[row_index(E)], dm_R_abs_colsum=np.abs(R).sum(axis=0))
        log["seconds_stages"] = time.time() - t0
        if a.f64:
            t0 = time.time()
            orc.set_inner_threads(a.threads, "f64")
            r64 = orc.mnem_em(D, sim["pert"], S, init, complete=sim["complete"], nthreads=1, acc="f64")
            out.update(f64_adj=r64["adj"].astype(np.uint8), f64_theta=r64["theta"].astype(np.int16), f64_ll=r64["ll"],
                       f64_n_iter=r64["n_iter"], f64_n_raw=r64["n_raw"])
            log.update(seconds_f64=time.time() - t0, f64_n_iter=r64["n_iter"].tolist(), f64_ll=r64["ll"].tolist(),
                       f64_phi_equal=[bool(np.array_equal(r64["adj"][s], ref["adj"][s])) for s in range(init.shape[0])])
        np.savez_compressed(os.path.join(HERE, "scale_%s.npz" % name), **out)
        log.update(n_iter=ref["n_iter"].tolist(), ll=ref["ll"].tolist(), min_margin=ref["min_margin"].tolist(),
                   min_accept_margin=ref["min_accept_margin"].tolist(), noise_modes=modes,
                   n_dec=ref["n_dec"].tolist(), n_dec_tight=ref["n_dec_tight"].tolist())
        print(json.dumps(log), flush=True)
        del D


if __name__ == "__main__":
    main()
