"""Full mnem() fits on the GPU against the CPU oracle at BASELINE scale (cfg3, cfg4, cfg5 shapes up to the full
30 x 2000 x 1 000 020 north-star size), through committed golden vectors.

The oracle needs the E x C matrix in host memory and minutes to hours per fit, so its results are computed once by
tests/golden/make_scale_golden.py (here, from the same seeded generator the device uses) and stored in
tests/golden/scale_*.npz.  Bar: phi, theta, iteration counts and the phi / theta change traces bit-exact; ll, L, mw to
1e-9 relative.  Every run also writes what it measured -- including the oracle's smallest greedy and accept margins
and, on a mismatch, what each side chose -- to gpurun_out/scale_parity_<case>.json (copied to profiles/).
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "golden"))
import mnem_b200 as mb  # noqa: E402
import make_scale_golden as msg  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-9
CASES = [c for c in msg.CASES if os.path.exists(os.path.join(HERE, "golden", "scale_%s.npz" % c))]


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def run_case(name, batch=0):
    g = np.load(os.path.join(HERE, "golden", "scale_%s.npz" % name))
    sim, init = msg.build_case(name)
    assert np.array_equal(init, g["init"]), "the seeded initial graphs changed: regenerate the golden"
    starts, k = init.shape[:2]
    with mb.Context.synthetic(sim) as cx:
        r = mb.em_run(cx, init, complete=sim["complete"], batch=batch)
    rep = dict(case=name, S=sim["S"], E=sim["E"], C=sim["C"], k=k, starts=starts, complete=bool(sim["complete"]),
               oracle_min_greedy_margin=g["min_margin"].tolist(), oracle_min_accept_margin=g["min_accept_margin"].tolist(),
               oracle_n_iter=g["n_iter"].tolist(), gpu_n_iter=r["n_iter"].tolist(), starts_detail=[])
    ok = True
    idx = g["probs_idx"]
    modes = [int(m) for m in g["modes"]]
    for s in range(starts):
        n = int(g["n_iter"][s])
        d = dict(start=s)
        d["phi_equal"] = bool(np.array_equal(r["adj"][s], g["adj"][s]))
        d["theta_equal"] = bool(np.array_equal(r["theta"][s], g["theta"][s]))
        d["phi_edges_differing"] = int((r["adj"][s] != g["adj"][s]).sum())
        d["theta_entries_differing"] = int((r["theta"][s] != g["theta"][s]).sum())
        d["n_iter_equal"] = bool(r["n_iter"][s] == n)
        m = min(n, int(r["n_iter"][s]))
        d["phievo_equal"] = bool(d["n_iter_equal"] and np.array_equal(r["phievo"][s], g["phievo"][s, :n]))
        d["thetaevo_equal"] = bool(d["n_iter_equal"] and np.array_equal(r["thetaevo"][s], g["thetaevo"][s, :n]))
        first_div = next((i for i in range(m) if r["phievo"][s][i] != g["phievo"][s, i] or
                          r["thetaevo"][s][i] != g["thetaevo"][s, i] or
                          abs(r["lls"][s][i] - g["lls"][s, i]) > RTOL * abs(g["lls"][s, i])), None)
        d["first_diverging_em_iteration"] = first_div
        d["ll_rel"] = _rel(r["ll"][s], g["ll"][s])
        d["lls_rel"] = _rel(r["lls"][s][:m], g["lls"][s, :m])
        d["probs_sample_rel"] = float(np.max(np.abs(r["probs"][s][:, idx] - g["probs_sample"][s]) /
                                             np.maximum(np.abs(g["probs_sample"][s]), 1.0)))
        d["probs_rowsum_rel"] = float(np.max(np.abs(r["probs"][s].sum(axis=1) - g["probs_rowsum"][s]) / g["probs_abs_rowsum"][s]))
        d["mw_rel"] = min(_rel(r["mw"][s], g["m%d_mw" % q][s]) for q in modes)
        exact = d["phi_equal"] and d["theta_equal"] and d["n_iter_equal"] and d["phievo_equal"] and d["thetaevo_equal"]
        close = max(d["ll_rel"], d["lls_rel"], d["probs_sample_rel"], d["probs_rowsum_rel"], d["mw_rel"]) <= RTOL
        d["green"] = bool(exact and close)
        ok = ok and d["green"]
        rep["starts_detail"].append(d)
    rep["cand_scored_equal"] = bool(r["stats"]["cand_scored"] == int(g["n_raw"].sum()))
    rep["gpu_cand_scored"], rep["oracle_cand_generated"] = int(r["stats"]["cand_scored"]), int(g["n_raw"].sum())
    top = g["ll"].max()
    rep["winner_gpu"], rep["winner_oracle"] = int(r["best"]), int(g["best"])
    rep["winner_ok"] = bool(abs(g["ll"][r["best"]] - top) <= RTOL * abs(top))
    rep["green"] = bool(ok and rep["winner_ok"] and rep["cand_scored_equal"])
    rep["ms_total_gpu"] = r["stats"]["ms_total"]
    out = os.path.join(HERE, "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "scale_parity_%s%s.json" % (name, "_b%d" % batch if batch else "")), "w") as f:
        json.dump(rep, f, indent=1)
    return rep


@pytest.mark.parametrize("name", CASES)
def test_fit_matches_oracle_at_scale(name):
    rep = run_case(name)
    bad = [d for d in rep["starts_detail"] if not d["green"]]
    assert rep["green"], json.dumps(dict(bad=bad, winner_ok=rep["winner_ok"], cand_scored_equal=rep["cand_scored_equal"]))


if __name__ == "__main__":
    for c in (sys.argv[1:] or CASES):
        print(json.dumps(run_case(c)))
