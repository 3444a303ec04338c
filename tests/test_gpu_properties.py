"""Size-independent properties of the CUDA path at sizes the oracle cannot finish in seconds (BASELINE configs 3-4
shapes: 20-24 S-genes, 1000 E-genes, 1e5 cells).  Run-to-run determinism, likelihood monotonicity, consistency of the
pieces with each other, linearity of the contractions, batch-size independence."""
import numpy as np
import pytest

import mnem_b200 as mb
import oracle as orc
from mnem_b200 import simdata

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    sim = simdata.make_config(dict(S=20, Egenes=50, C=100000, mw=[0.2] * 5, k=5, starts=6, uninform=0, complete=True),
                              seed=1003)
    cx = mb.Context.synthetic(sim)
    init = simdata.init_phi(sim["S"], 5, 6, seed=77)
    yield sim, cx, init
    cx.close()


def test_em_is_deterministic_and_batch_independent(big):
    sim, cx, init = big
    a = mb.em_run(cx, init, complete=True, batch=4)
    b = mb.em_run(cx, init, complete=True, batch=4)
    c = mb.em_run(cx, init, complete=True, batch=1)
    d = mb.em_run(cx, init, complete=True, batch=6)           # 30 (start, component) pairs per pass
    for key in ("adj", "theta", "ll", "mw", "n_iter", "probs"):
        assert np.array_equal(a[key], b[key]), key            # bit-reproducible run to run
        assert np.array_equal(a[key], c[key]), key            # sharing passes between starts changes nothing
        assert np.array_equal(a[key], d[key]), key
    assert a["best"] == b["best"] == c["best"]
    for s in range(init.shape[0]):
        lls = a["lls"][s]
        assert len(lls) == a["n_iter"][s] >= 2
        assert (np.diff(lls) >= 0).all()                      # accepted likelihood never decreases (R/mnems.r:2104-2108)
        assert a["ll"][s] == lls.max()
        np.testing.assert_allclose(a["mw"][s].sum(), 1.0, rtol=1e-12)
        # the loop ends by the stop rule (no edge / theta change, R/mnems.r:2138-2141) or because the last E-step was
        # not accepted (ll unchanged, :1955-1958)
        stopped = a["phievo"][s][-1] == 0 and a["thetaevo"][s][-1] == 0
        assert stopped or lls[-1] == lls[-2]
    assert a["best"] == max(i for i in range(len(a["ll"])) if a["ll"][i] == a["ll"].max())


def test_shared_start_queue_gives_the_same_fits(big):
    """Two callers draining one queue (what the ranks of an N-GPU job do) fit every start exactly once, and every fit is
    bit-identical to the single-call fit: a start's result does not depend on which starts share its passes."""
    from mnem_b200 import dist as mdist
    sim, cx, init = big
    full = mb.em_run(cx, init, complete=True, batch=2)
    q = mdist.StartQueue(init.shape[0])
    first = {"n": 0}

    def two_only():                       # caller A stops asking after two starts ...
        if first["n"] >= 2:
            return -1
        first["n"] += 1
        return q()
    a = mb.em_run(cx, init, complete=True, batch=2, next_start=two_only)
    b = mb.em_run(cx, init, complete=True, batch=2, next_start=q)      # ... caller B drains the rest
    ida = [s for s in range(init.shape[0]) if a["n_iter"][s] >= 0]
    idb = [s for s in range(init.shape[0]) if b["n_iter"][s] >= 0]
    assert sorted(ida + idb) == list(range(init.shape[0])) and len(ida) == 2
    assert np.isneginf(a["ll"][idb]).all() and np.isneginf(b["ll"][ida]).all()
    for part, ids in ((a, ida), (b, idb)):
        for key in ("adj", "theta", "ll", "mw", "n_iter", "probs"):
            assert np.array_equal(part[key][ids], full[key][ids]), key
    lls = np.where(a["n_iter"] >= 0, a["ll"], b["ll"])
    assert mdist.pick_winner(lls) == full["best"]


def test_em_pieces_are_consistent_at_scale(big):
    """The winner's stored probs must be reproduced by getProbs from its (phi, theta), and its graphs must be fixed
    points of transitive reduction; the planted mixture is recovered well."""
    sim, cx, init = big
    r = mb.em_run(cx, init, complete=True)
    b = r["best"]
    res = [dict(adj=r["adj"][b][i], subtopo=r["theta"][b][i]) for i in range(5)]
    pr = mb.getProbs(r["probs"][b], 5, cx, res, mw=r["mw"][b], complete=True)
    np.testing.assert_array_equal(pr["probs"], r["probs"][b])
    for i in range(5):
        g = r["adj"][b][i]
        assert np.array_equal(mb.transitive_reduction(g), g)
        assert (np.diag(g) == 0).all()
    post = mb.getAffinity(r["probs"][b], mw=r["mw"][b], complete=True)
    hard = post.argmax(0)
    # components are recovered up to permutation: every true component maps mostly onto one fitted component
    purity = 0
    for t in range(5):
        cnt = np.bincount(hard[sim["comp"] == t], minlength=5)
        purity += cnt.max()
    assert purity / sim["C"] > 0.35          # well above the 1/k of a random assignment


def test_contractions_are_linear_and_match_a_cpu_subsample(big):
    sim, cx, init = big
    rng = np.random.default_rng(0)
    w1, w2 = rng.random(sim["C"]), rng.random(sim["C"])
    R1, R2, R12 = mb.doMean(cx, w1), mb.doMean(cx, w2), mb.doMean(cx, 2.0 * w1 + w2)
    np.testing.assert_allclose(R12, 2.0 * R1 + R2, rtol=1e-11, atol=1e-8)
    # indicator weights of the first 3000 cells: equals the oracle on that sub-matrix
    n = 3000
    D = cx.read_data(0, n)
    assert np.array_equal(D, simdata.host_data(sim, 0, n))
    ind = np.zeros(sim["C"])
    ind[:n] = 1.0
    np.testing.assert_allclose(mb.doMean(cx, ind), orc.do_mean(D, sim["pert"][:n], sim["S"]), rtol=1e-11, atol=1e-9)
    # E-step on the full matrix, checked on the same 3000 cells
    phi = init[0]
    theta = rng.integers(0, sim["S"] + 1, (5, sim["E"])).astype(np.int32)
    res = [dict(adj=phi[i], subtopo=theta[i]) for i in range(5)]
    got = mb.getProbs(np.zeros((5, sim["C"])), 5, cx, res, complete=True)
    want = orc.get_probs(D, sim["pert"][:n], sim["S"], phi, theta, np.zeros((5, n)), None, complete=True)
    np.testing.assert_allclose(got["probs"][:, :n], want["probs"], rtol=1e-12, atol=1e-12)


def test_greedy_search_at_config5_shape():
    """S = 30, E = 2000: the device search against the oracle on one planted problem (seconds on the CPU)."""
    rng = np.random.default_rng(5)
    S, E = 30, 2000
    truth = orc.mytc(np.triu((rng.random((S, S)) < 0.25).astype(np.uint8), 1))
    p = rng.permutation(S)
    truth = truth[np.ix_(p, p)]
    th = rng.integers(1, S + 1, E)
    R = np.where(truth[:, th - 1].T > 0, 1.0, -1.0) * 4 + rng.normal(size=(E, S)) * 3
    want = orc.nem_greedy(R, None)
    got = mb.nem_greedy(R, None)
    assert want["min_margin"] > 1e-13
    assert np.array_equal(got["adj"], want["adj"]) and np.array_equal(got["subtopo"], want["theta"])
    assert got["n_iter"] == want["n_iter"] and got["n_scored"] == want["n_raw"]
    assert abs(got["score"] - want["score"]) <= 1e-12 * abs(want["score"])


def test_config5_full_size_properties():
    """BASELINE config 5 at full size (30 S-genes x 2000 E-genes x 1 000 020 cells, k = 5; 2 x 16 GB in HBM), 3 starts:
    run-to-run determinism, monotone accepted likelihood, the winner's probs reproduced by getProbs from its
    (phi, theta), doMean linearity, and the E-step checked against the oracle on a 2000-cell slice."""
    sim = simdata.make_config(5)
    init = simdata.init_phi(sim["S"], sim["k"], 3, seed=2005)
    with mb.Context.synthetic(sim) as cx:
        a = mb.em_run(cx, init, complete=True)
        b = mb.em_run(cx, init, complete=True, batch=2)
        for key in ("adj", "theta", "ll", "mw", "n_iter", "probs"):
            assert np.array_equal(a[key], b[key]), key
        for s in range(3):
            assert (np.diff(a["lls"][s]) >= 0).all() and a["ll"][s] == a["lls"][s].max()
            for g in a["adj"][s]:
                assert np.array_equal(mb.transitive_reduction(g), g)
        w = a["best"]
        res = [dict(adj=a["adj"][w][i], subtopo=a["theta"][w][i]) for i in range(5)]
        pr = mb.getProbs(a["probs"][w], 5, cx, res, mw=a["mw"][w], complete=True)
        assert np.array_equal(pr["probs"], a["probs"][w])
        n = 2000
        D = cx.read_data(500000, n)
        assert np.array_equal(D, simdata.host_data(sim, 500000, n))
        want = orc.get_probs(D, sim["pert"][500000:500000 + n], sim["S"], a["adj"][w], a["theta"][w], np.zeros((5, n)), None,
                             complete=True)
        np.testing.assert_allclose(a["probs"][w][:, 500000:500000 + n], want["probs"], rtol=1e-12, atol=1e-12)
        rng = np.random.default_rng(1)
        w1, w2 = rng.random(sim["C"]), rng.random(sim["C"])
        R1, R2, R12 = mb.doMean(cx, w1), mb.doMean(cx, w2), mb.doMean(cx, w1 + 3.0 * w2)
        np.testing.assert_allclose(R12, R1 + 3.0 * R2, rtol=1e-10, atol=1e-7)
        # the planted mixture is recovered: hard assignments are far purer than chance
        post = mb.getAffinity(a["probs"][w], mw=a["mw"][w], complete=True)
        hard = post.argmax(0)
        purity = sum(np.bincount(hard[sim["comp"] == t], minlength=5).max() for t in range(5)) / sim["C"]
        assert purity > 0.35


def test_config4_shape_model_selection():
    """BASELINE config 4 shape (24 S-genes, 41 E-genes each + 16 uninformative = 1000, three planted components), fewer
    cells: mnemk() over k = 1..4 with getIC() (R/mnems.r:1477-1529, 1560-1574) -- the criterion is finite for every k,
    its arithmetic matches the formula recomputed from the fits, and the chosen k minimises it."""
    cfg = dict(simdata.CONFIGS[4])
    cfg["C"] = 30000
    sim = simdata.make_config(cfg, seed=1004)
    with mb.Context.synthetic(sim) as cx:
        out = mb.mnemk(cx, ks=(1, 2, 3, 4), starts=2, complete=True, seed=3)
    assert set(out["ics"]) == {1, 2, 3, 4} and all(np.isfinite(v) for v in out["ics"].values())
    assert out["k"] == min(out["ics"], key=lambda q: out["ics"][q])
    assert out["lls"][3] > out["lls"][1]                           # three planted components beat a single network
    fit = out["best"]
    k = out["k"]
    fpar = sum(int((orc.mytc(c["phi"]) - np.eye(24, dtype=int) != 0).sum()) for c in fit["comp"]) + k * sim["E"] + (k - 1)
    ic = np.log(sim["C"]) * fpar - 2 * np.max(fit["ll"]) * np.log(2)
    assert abs(ic - out["ics"][k]) <= 1e-9 * abs(ic)


@pytest.mark.parametrize("S,E,n", [(1, 5, 2), (2, 40, 3), (7, 130, 9), (13, 33, 5), (20, 1000, 12), (30, 2000, 10), (32, 700, 6)])
def test_device_resident_greedy_loop_equals_the_host_driven_one(S, E, n, monkeypatch):
    """The greedy loop runs inside one persistent kernel (search_persistent_kernel: a queue of (search, tiles) tasks, the
    CTA that scores the last tile of an iteration finishes it and publishes the next).  Every number must be produced by
    the same arithmetic as in the host-driven schedule (one launch per step, MNEMCU_SEARCH=host): graphs, theta, scores,
    traces and counts are compared bit for bit, from empty and from random non-closed starts (exact score ties included:
    duplicated and zero columns)."""
    rng = np.random.default_rng(100 * S + n)
    R = rng.normal(size=(n, E, S)) * 3
    if S > 2:
        R[:, :, 1] = R[:, :, 0]                                   # duplicated S-gene: structural ties
        R[0, :, 2] = 0.0
    starts = np.stack([np.triu((rng.random((S, S)) < 0.15).astype(np.uint8), 1) for _ in range(n)])
    for st in (None, starts):
        monkeypatch.setenv("MNEMCU_SEARCH", "host")
        a = mb.nem_greedy(R, st)
        monkeypatch.setenv("MNEMCU_SEARCH", "device")
        b = mb.nem_greedy(R, st)
        for key in ("adj", "subtopo", "score", "max_trace", "n_iter", "n_scored"):
            assert np.array_equal(a[key], b[key]), key
        assert b["n_iter"].max() > 0 or S == 1


def test_em_fit_is_the_same_on_both_greedy_schedules(monkeypatch):
    sim = simdata.make_config(dict(S=9, Egenes=12, C=4000, mw=[0.3, 0.3, 0.4], k=3, starts=5, uninform=3), seed=77)
    init = simdata.init_phi(9, 3, 5, seed=78)
    with mb.Context.synthetic(sim) as cx:
        monkeypatch.setenv("MNEMCU_SEARCH", "host")
        a = mb.em_run(cx, init, complete=True)
        monkeypatch.setenv("MNEMCU_SEARCH", "device")
        b = mb.em_run(cx, init, complete=True)
    for key in ("adj", "theta", "ll", "mw", "n_iter", "probs", "best"):
        assert np.array_equal(a[key], b[key]), key
    assert a["stats"]["cand_scored"] == b["stats"]["cand_scored"] and a["stats"]["greedy_iters"] == b["stats"]["greedy_iters"]
