"""bench.py's CPU baseline is a cost model assembled from bounded component timings of the oracle (the full workload would
take hours on a core).  On a problem small enough to run to convergence the model must reproduce the real run."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import oracle as orc  # noqa: E402
from mnem_b200 import simdata  # noqa: E402


def test_cpu_cost_model_matches_a_full_oracle_run():
    sim = simdata.make_config(dict(S=12, Egenes=30, C=24000, mw=[1 / 3] * 3, k=3, starts=2, uninform=0, complete=True), seed=5)
    init = simdata.init_phi(12, 3, 2, seed=1)
    D = simdata.host_data(sim)
    t0 = time.time()
    ref = orc.mnem_em(D, sim["pert"], sim["S"], init[:1], complete=True, nthreads=1, faithful_cost=True)
    actual = ref["n_iter"].sum() / (time.time() - t0)
    model = bench.cpu_extrapolated(sim, init, 1, 3000, float(ref["n_iter"].mean()))
    assert 0.4 * actual <= model["value"] <= 2.5 * actual, (actual, model["value"])   # wall-clock timing on a shared box: generous
    assert model["cells"] == 3000 and "scaled by 24000/3000" in model["sample"]
