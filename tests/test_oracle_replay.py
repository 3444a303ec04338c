"""The decision replay of the oracle (oracle/mnem_oracle.c, orc_advice), on the CPU.

The tape here is written by the oracle built with DOUBLE accumulators (R where long double is 64 bits, e.g. arm64
macOS) and replayed by the build with LONG-DOUBLE accumulators (R on x86-64): two correct implementations of the
reference that round differently -- the situation of a device fit (tests/test_gpu_scale_parity.py).  Left alone the
two builds return different graphs at this size; replayed, they agree, and every overridden decision has a margin at
rounding level.  A tape that takes a genuinely worse move must be rejected."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle as orc  # noqa: E402
from mnem_b200 import simdata  # noqa: E402
from mnem_b200.api import parse_tape  # noqa: E402


def _problem(C=8000, seed=1003):
    cfg = dict(simdata.CONFIGS[3])
    cfg.update(C=C, starts=1, Egenes=20)
    sim = simdata.make_config(cfg, seed=seed)
    init = simdata.init_phi(sim["S"], sim["k"], 1, seed=2003)
    D = simdata.host_data(sim)
    return sim, init, D


def test_replay_reconciles_the_two_accumulator_widths():
    sim, init, D = _problem()
    S, E, k = sim["S"], sim["E"], sim["k"]
    n = min(8, len(os.sched_getaffinity(0)))
    for acc in ("ld", "f64"):
        orc.set_inner_threads(n, acc)
    try:
        rec = orc.record_start(D, sim["pert"], S, init[0], complete=sim["complete"], acc="f64")
        tape = parse_tape(rec["words"], k, E, S)[0]
        assert tape["n_iter"] == rec["n_iter"] and len(tape["theta"]) == rec["n_iter"]
        rp = orc.replay_start(D, sim["pert"], S, init[0], tape, complete=sim["complete"])
        assert not rp["diverged"], rp["first_hard"]
        assert sum(v["hard"] for v in rp["report"].values()) == 0
        assert np.array_equal(rp["adj"], rec["adj"]) and np.array_equal(rp["theta"], rec["theta"])
        assert rp["n_iter"] == rec["n_iter"]
        assert abs(rp["ll"] - rec["ll"]) <= 1e-9 * abs(rec["ll"])
        assert np.max(np.abs(rp["lls"] - rec["lls"]) / np.abs(rec["lls"])) <= 1e-9
        assert max(v["max_forced"] for q, v in rp["report"].items() if q in ("greedy_move", "greedy_stop", "restart_pick")) <= 1e-13
        assert rp["report"]["greedy_move"]["checked"] > 100
        # a tape of the SAME build is followed without a single override
        rec2 = orc.record_start(D, sim["pert"], S, init[0], complete=sim["complete"], acc="ld")
        rp2 = orc.replay_start(D, sim["pert"], S, init[0], parse_tape(rec2["words"], k, E, S)[0], complete=sim["complete"])
        assert not rp2["diverged"] and sum(v["forced"] for v in rp2["report"].values()) == 0
        assert np.array_equal(rp2["adj"], rec2["adj"]) and rp2["ll"] == rec2["ll"]
    finally:
        for acc in ("ld", "f64"):
            orc.set_inner_threads(1, acc)


def test_replay_rejects_a_wrong_move():
    sim, init, D = _problem(C=3000)
    S, E, k = sim["S"], sim["E"], sim["k"]
    rec = orc.record_start(D, sim["pert"], S, init[0], complete=sim["complete"], acc="ld")
    tape = parse_tape(rec["words"], k, E, S)[0]
    # first search with at least two moves: replace its first move by the graph after a DIFFERENT legal move
    key = next(q for q in sorted(tape["searches"]) if len(tape["searches"][q]) >= 2)
    first = tape["searches"][key][0].copy()
    assert key[2] == 0                               # restart 0 climbs from the empty graph (R/mnems.r:2042)
    models = orc.neighbours(np.eye(S, dtype=np.uint8))[0]

    def rows_of(gph):
        return np.array([sum(int(gph[s, j] != 0) << j for j in range(S)) for s in range(S)], dtype=np.uint32)
    other = next(rows_of(gph) for gph in models[::-1] if not np.array_equal(rows_of(gph), first))
    bad = dict(tape)
    bad["searches"] = dict(tape["searches"])
    bad["searches"][key] = np.vstack([other[None], tape["searches"][key][1:]])
    rp = orc.replay_start(D, sim["pert"], S, init[0], bad, complete=sim["complete"])
    assert rp["diverged"] and sum(v["hard"] for v in rp["report"].values()) >= 1
