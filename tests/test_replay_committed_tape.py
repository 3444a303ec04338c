"""A device fit at BASELINE scale, checked WITHOUT a GPU.

tests/golden/gpu_tape_cfg3_full_s0.npz was recorded on a B200 by scripts/save_tape.py: the decision tape of the fit of
start 0 of BASELINE config 3 at full size (20 S-genes x 1000 E-genes x 100 000 cells, k = 5) plus what the device
returned.  Here the same synthetic matrix is regenerated on the CPU (oracle/simgen.c) and the oracle replays the tape
in its own arithmetic: every decision equal or within 1e-13, no genuine disagreement, and the
replayed fit equals the device's -- graphs, theta, iteration count and change traces bit for bit, likelihoods to 1e-9."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_scale_golden as msg  # noqa: E402
import oracle as orc  # noqa: E402
from mnem_b200.api import parse_tape  # noqa: E402

TAPE = os.path.join(HERE, "golden", "gpu_tape_cfg3_full_s0.npz")


@pytest.mark.skipif(not os.path.exists(TAPE), reason="no committed device tape")
def test_committed_device_tape_replays_through_the_oracle():
    g = np.load(TAPE)
    name, start = str(g["case"]), int(g["start"])
    sim, init = msg.build_case(name)
    assert np.array_equal(init[start], g["init"])
    n = max(1, min(16, len(os.sched_getaffinity(0))))
    D = orc.simgen(sim, nthreads=n)
    orc.set_inner_threads(n)
    try:
        tape = parse_tape(g["tape_words"], sim["k"], sim["E"], sim["S"])[0]
        o = orc.replay_start(D, sim["pert"], sim["S"], init[start], tape, complete=sim["complete"])
    finally:
        orc.set_inner_threads(1)
    assert not o["diverged"], o["first_hard"]
    assert sum(v["hard"] for v in o["report"].values()) == 0
    assert max(v["max_forced"] for q, v in o["report"].items() if q in ("greedy_move", "greedy_stop", "restart_pick")) <= 1e-13
    assert o["report"]["greedy_move"]["checked"] > 500
    assert np.array_equal(o["adj"], g["adj"]) and np.array_equal(o["theta"], g["theta"])
    assert o["n_iter"] == int(g["n_iter"])
    assert np.array_equal(o["phievo"], g["phievo"]) and np.array_equal(o["thetaevo"], g["thetaevo"])
    assert abs(o["ll"] - float(g["ll"])) <= 1e-9 * abs(float(g["ll"]))
    assert np.max(np.abs(o["lls"] - g["lls"]) / np.abs(g["lls"])) <= 1e-9
    assert np.allclose(o["mw"], g["mw"], rtol=1e-9, atol=0)
    ps = o["probs"][:, g["probs_idx"]]
    assert np.max(np.abs(ps - g["probs_sample"]) / np.maximum(np.abs(g["probs_sample"]), 1.0)) <= 1e-9
