"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (north star): graphs, theta, closures and every other integer output bit-exact; log-likelihoods,
responsibilities, mixture weights and scores within 1e-9 relative (tolerances written at each assert; most
comparisons are far tighter because both sides use fixed summation orders).
"""
import os

import numpy as np
import pytest

import mnem_b200 as mb
import oracle as orc
from mnem_b200 import simdata

pytestmark = pytest.mark.gpu
RTOL = 1e-9
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "app_rda_golden.npz"))
KEYS = sorted(k[:-6] for k in G.files if k.endswith("_probs"))


def rand_graph(rng, S, p=0.25, dag=False):
    a = (rng.random((S, S)) < p).astype(np.uint8)
    if dag:
        a = np.triu(a, 1)
    np.fill_diagonal(a, 0)
    return a


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


# ---- src/mm.cpp -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S", [1, 2, 3, 7, 13, 24, 31, 32])
def test_mm_cpp_graph_routines(S):
    rng = np.random.default_rng(S)
    gs = np.stack([rand_graph(rng, S, rng.uniform(0.02, 0.5), dag=i % 2 == 0) for i in range(24)])
    got = mb.transClose_W(gs)
    for g, o in zip(gs, got):
        assert (o == orc.trans_close_w(g)).all()
    got = mb.transitive_closure(gs)
    red = mb.transitive_reduction(gs)
    for g, o, r in zip(gs, got, red):
        assert (o == orc.mytc(g)).all()
        assert (r == orc.transitive_reduction(g)).all()
    u = rng.integers(1, S + 1, len(gs))
    v = rng.integers(1, S + 1, len(gs))
    ins = mb.transClose_Ins(gs, u, v)
    dele = mb.transClose_Del(gs, u, v)
    for i, g in enumerate(gs):
        assert (ins[i] == orc.trans_close_ins(g, u[i], v[i])).all()
        assert (dele[i] == orc.trans_close_del(g, u[i], v[i])).all()


def test_closure_roxygen_example():
    g = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], dtype=float).reshape(3, 3, order="F")     # R/mnems.r:547
    assert (mb.transitive_closure(g) == np.triu(np.ones((3, 3)))).all()
    assert (mb.transitive_closure(g, 1, 2) == orc.mytc(g, 1, 2)).all()


def test_maxcol_and_matmul():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(257, 9))
    x[5] = x[5, 0]                       # ties -> first
    x[6, :] = -np.inf
    assert (mb.maxCol_row(x) == orc.max_col_row(x)).all()
    A, B = rng.normal(size=(131, 17)), rng.normal(size=(17, 5))
    assert np.array_equal(mb.eigenMapMatMult(A, B), orc.matmul(A, B))   # same ascending mul/add order: bit-exact


@pytest.mark.parametrize("S", [2, 3, 5, 9, 16, 30])
def test_neighbours(S):
    rng = np.random.default_rng(300 + S)
    for t in range(8):
        P = rand_graph(rng, S, rng.uniform(0.03, 0.4), dag=t % 2 == 0)
        if t % 3 == 0:
            P = orc.mytc(P)
        np.fill_diagonal(P, 1)
        got, kind = mb.neighbours(P)
        want, wkind, nraw = orc.neighbours(P)
        assert len(got) == nraw
        uniq, ukind = [], []
        for m, kd in zip(got, kind):                        # unique(): keep first, R/mnems.r:222
            if not any((m == q).all() for q in uniq):
                uniq.append(m)
                ukind.append(kd)
        assert len(uniq) == len(want)
        assert all((a == b).all() for a, b in zip(uniq, want))
        assert list(ukind) == list(wkind)


# ---- responsibilities / likelihood ------------------------------------------------------------------------------
@pytest.mark.parametrize("key", KEYS)
def test_getll_app_rda_known_answers(key):
    """The reference's own fitted objects: ll must equal getLL(probs, mw) (data/app.rda)."""
    probs, mw, ll = G[key + "_probs"], G[key + "_mw"], G[key + "_ll"]
    got = mb.getLL(probs, mw=mw)
    assert abs(got - ll.max()) <= 1e-12 * abs(ll.max())
    g = mb.getAffinity(probs, mw=mw)
    close(g, orc.get_affinity(probs, mw), rtol=1e-12)
    if probs.shape[0] > 1:                                   # inst/unitTests/test_hardClustering.R
        hard = mb.getAffinity(probs, affinity=1, mw=mw)
        assert (hard == orc.get_affinity(probs, mw, affinity=1)).all()
        assert (hard == (g == g.max(0, keepdims=True))).all()


@pytest.mark.parametrize("k,complete,logtype", [(1, False, 2), (2, False, 2), (3, True, 2), (5, True, 2), (4, False, np.e),
                                                (3, True, 10), (8, True, 2), (9, False, 2), (13, True, 2), (32, True, 2),
                                                (32, False, 2)])
def test_affinity_and_ll_random(k, complete, logtype):
    rng = np.random.default_rng(k)
    L = rng.normal(size=(k, 3001)) * 40
    mw = rng.random(k)
    mw /= mw.sum()
    close(mb.getAffinity(L, mw=mw, complete=complete, logtype=logtype),
          orc.get_affinity(L, mw, complete=complete, logtype=logtype), rtol=1e-11, atol=1e-300)
    got, want = mb.getLL(L, mw=mw, complete=complete, logtype=logtype), orc.get_ll(L, mw, complete=complete, logtype=logtype)
    assert abs(got - want) <= 1e-11 * abs(want)
    close(mb.getAffinity(-np.abs(L), complete=complete), orc.get_affinity(-np.abs(L), None, complete=complete), rtol=1e-11,
          atol=1e-300)


def test_overflow_semantics_q9():
    """complete = FALSE: 2^L overflows for L >= 1024 -> ll = Inf, gamma NaN -> 0."""
    L = np.array([[1100.0, 3.0, -200.0], [1100.0, 1.0, -210.0]])
    assert mb.getLL(L) == orc.get_ll(L) == np.inf
    Lu = np.array([[1100.0, -2000.0], [1100.0, -2100.0]])           # Inf + log(0): NaN on both sides
    assert np.isnan(mb.getLL(Lu)) and np.isnan(orc.get_ll(Lu))
    assert np.array_equal(mb.getAffinity(L), orc.get_affinity(L))
    close(mb.getAffinity(L, complete=True), orc.get_affinity(L, complete=True), rtol=1e-12)


# ---- M-step statistics, scoring, greedy search -------------------------------------------------------------------
@pytest.mark.parametrize("S,E,Cn", [(3, 1, 7), (5, 10, 1000), (10, 101, 5003), (30, 257, 4000)])
def test_do_mean(S, E, Cn):
    rng = np.random.default_rng(S * E)
    D = rng.normal(size=(E, Cn))
    pert = rng.integers(1, S + 1, Cn).astype(np.int32)
    pert[:S] = np.arange(1, S + 1)
    w = rng.random((7, Cn))
    w[rng.random((7, Cn)) < 0.05] = 0.0
    with mb.Context(D, pert) as cx:
        R = mb.doMean(cx, w)
        R0 = mb.doMean(cx)
    for j in range(7):
        close(R[j], orc.do_mean(D, pert, S, w[j]), rtol=1e-11, atol=1e-11)
    close(R0, orc.do_mean(D, pert, S), rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("S,E", [(2, 3), (5, 10), (10, 100), (20, 333), (30, 1000), (32, 129)])
def test_score_adj(S, E):
    rng = np.random.default_rng(S + E)
    R = rng.normal(size=(E, S)) * 3
    adjs = np.stack([rand_graph(rng, S, rng.uniform(0.02, 0.5)) for _ in range(40)])
    for tc in (True, False):
        got = mb.scoreAdj(R, adjs, trans_close=tc)
        for i, a in enumerate(adjs):
            sc, th, W = orc.score_adj(R, a, trans_close=tc)
            assert np.array_equal(got["subweights"][i], W)          # same ascending S-gene sums: bit-exact
            assert np.array_equal(got["subtopo"][i], th)
            assert abs(got["score"][i] - sc) <= 1e-12 * abs(sc) + 1e-300


def planted_R(rng, S, E, snr=4.0):
    truth = orc.mytc(rand_graph(rng, S, 0.3, dag=True))
    truth = truth[np.ix_(rng.permutation(S), rng.permutation(S))]
    theta = rng.integers(1, S + 1, E)
    return np.where(truth[:, theta - 1].T > 0, 1.0, -1.0) * snr + rng.normal(size=(E, S)) * 3


@pytest.mark.parametrize("S,E,n", [(2, 5, 4), (3, 10, 6), (5, 20, 8), (8, 60, 6), (10, 100, 6), (16, 130, 4), (20, 300, 2)])
def test_nem_greedy(S, E, n):
    rng = np.random.default_rng(1000 + S)
    R = np.stack([planted_R(rng, S, E) for _ in range(n)])
    starts = np.stack([orc.transitive_reduction(rand_graph(rng, S, 0.15, dag=i % 2 == 0)) if i % 3 else np.zeros((S, S))
                       for i in range(n)])
    got = mb.nem_greedy(R, starts)
    got0 = mb.nem_greedy(R, None)
    for i in range(n):
        for g, st in ((got, starts[i]), (got0, None)):
            want = orc.nem_greedy(R[i], st)
            assert want["min_margin"] > 1e-11, "test case has a near-tie; pick another seed"
            assert np.array_equal(g["adj"][i], want["adj"])
            assert np.array_equal(g["subtopo"][i], want["theta"])
            assert g["n_iter"][i] == want["n_iter"]
            assert g["n_scored"][i] == want["n_raw"]
            assert abs(g["score"][i] - want["score"]) <= 1e-12 * abs(want["score"])
            assert abs(g["max_trace"][i] - want["max_trace"]) <= 1e-12 * abs(want["max_trace"])


def test_nem_greedy_structural_ties():
    """Exact score ties (duplicate E-gene rows, S-genes without effects): first maximum / sparser graph must win
    exactly as in R/mnems.r:236, 279-283."""
    rng = np.random.default_rng(77)
    S, E = 6, 24
    R = planted_R(rng, S, E)
    R[:, 4] = R[:, 3]                   # two indistinguishable S-genes
    R[:, 5] = -1.0                      # an S-gene no E-gene likes
    R = np.round(R * 4) / 4             # coarse values: many exact ties
    want = orc.nem_greedy(R, None)
    got = mb.nem_greedy(R, None)
    assert np.array_equal(got["adj"], want["adj"]) and np.array_equal(got["subtopo"], want["theta"])
    assert got["n_iter"] == want["n_iter"] and abs(got["score"] - want["score"]) <= 1e-12 * abs(want["score"])


# ---- E-step ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,E,Cn,k,complete", [(3, 4, 50, 2, False), (5, 10, 1000, 2, False), (7, 33, 2100, 3, True),
                                               (10, 100, 3001, 4, False), (30, 300, 5000, 5, True), (6, 20, 777, 1, False)])
def test_get_probs_theta_given(S, E, Cn, k, complete):
    rng = np.random.default_rng(S * k)
    D = rng.normal(size=(E, Cn))
    pert = rng.integers(1, S + 1, Cn).astype(np.int32)
    res = [dict(adj=orc.transitive_reduction(rand_graph(rng, S, 0.3)), subtopo=rng.integers(0, S + 1, E).astype(np.int32))
           for _ in range(k)]
    L_in = rng.normal(size=(k, Cn))
    mw = rng.random(k)
    mw /= mw.sum()
    got = mb.getProbs(L_in, k, D, res, pert=pert, mw=mw, complete=complete)
    want = orc.get_probs(D, pert, S, np.stack([r["adj"] for r in res]), np.stack([r["subtopo"] for r in res]), L_in, mw,
                         complete=complete)
    close(got["probs"], want["probs"], rtol=1e-11, atol=1e-11)
    close(got["mw"], want["mw"], rtol=1e-10)
    assert abs(got["ll"] - want["ll"]) <= RTOL * abs(want["ll"])


@pytest.mark.parametrize("S,E,Cn,k", [(3, 6, 90, 2), (5, 10, 1000, 2), (8, 40, 2500, 3), (20, 120, 4000, 5)])
def test_get_probs_theta_free(S, E, Cn, k):
    """First E-step of a start: theta = argmax with the null column last (R/mnems_low.r:617-619)."""
    sim = simdata.make_config(dict(S=S, Egenes=E // S, C=Cn, mw=[1.0 / k] * k, k=k, starts=1, uninform=E % S), seed=S)
    D = simdata.host_data(sim) + np.random.default_rng(S).normal(size=(sim["E"], sim["C"])) * 0.25
    phi = simdata.init_phi(S, k, 1, seed=5)[0]
    res = [dict(adj=p) for p in phi]
    L0 = np.zeros((k, sim["C"]))
    got = mb.getProbs(L0, k, D, res, pert=sim["pert"])
    want = orc.get_probs(D, sim["pert"], S, phi, None, L0, None)
    assert np.array_equal(got["theta_used"], want["theta_used"])
    close(got["probs"], want["probs"], rtol=1e-10, atol=1e-10)
    close(got["mw"], want["mw"], rtol=RTOL)
    assert abs(got["ll"] - want["ll"]) <= RTOL * abs(want["ll"])


# ---- the EM loop -------------------------------------------------------------------------------------------------
def check_em(sim, D, init, complete, batch=0, gpu_pert=None, affinity=0):
    """gpu_pert: what the CUDA side is given when it differs from the oracle's encoding sim["pert"] (the Rho path:
    "_"-joined labels for the GPU, -mask integers for the oracle)."""
    refs = [orc.mnem_em(D, sim["pert"], sim["S"], init, complete=complete, nthreads=8, noise_mode=m, affinity=affinity)
            for m in (0, 1, 2)]
    ref = refs[0]
    res = mb.mnem(D, sim["pert"] if gpu_pert is None else gpu_pert, phi=init, complete=complete, batch=batch,
                  affinity=affinity)
    starts = init.shape[0]
    # Decision margins: a greedy / restart decision closer than 1e-13 relative is rounding noise
    # in ANY implementation -- such a case cannot pin anything, pick another seed.
    assert ref["min_margin"].min() > 1e-13, "a greedy decision in this case is a near-tie; pick another seed"
    for s in range(starts):
        lim = res["limits"][s]
        assert np.array_equal(lim["adj"], ref["adj"][s]), "phi of start %d" % s
        assert np.array_equal(lim["theta"], ref["theta"][s]), "theta of start %d" % s
        assert res["n_iter"][s] == ref["n_iter"][s]
        assert abs(lim["ll"] - ref["ll"][s]) <= RTOL * abs(ref["ll"][s])
        close(lim["probs"], ref["probs"][s], rtol=RTOL, atol=1e-9)
        close(lim["lls"], ref["lls"][s], rtol=RTOL)
        assert np.array_equal(lim["phievo"], ref["phievo"][s]) and np.array_equal(lim["thetaevo"], ref["thetaevo"][s])
        close(lim["mwevo"], ref["mwevo"][s], rtol=1e-6, atol=1e-12)
        # limits$mw is the mixture weight at the last iteration whose ll was STRICTLY greater (R/mnems.r:1960-1962,
        # 2104, 2144-2146).  Once the graphs stop changing that comparison is decided by the last bit of ll (the
        # oracle reports the margin), so both outcomes are legitimate: accept either, each to 1e-9.
        cands = [ref] if ref["min_accept_margin"][s] > 1e-12 else refs
        assert any(np.allclose(lim["mw"], r["mw"][s], rtol=RTOL, atol=0) for r in cands), (lim["mw"], [r["mw"][s] for r in cands])
    # winner: last maximal ll (R/mnems.r:2226).  Starts that reach the same optimum differ in the last bits of ll
    # (SURVEY Q12), so the GPU's winner must be one of the oracle's maximal starts up to 1e-9.
    top = ref["ll"].max()
    assert abs(ref["ll"][res["best_start"]] - top) <= RTOL * abs(top)
    assert res["stats"]["em_iters"] == ref["n_iter"].sum()
    assert res["stats"]["cand_scored"] == ref["n_raw"].sum()
    b = res["best_start"]
    lams = [orc.final_mw(r["probs"][b], r["mw"][b], complete=complete, affinity=affinity) for r in refs]
    assert any(np.allclose(res["mw"], lam, rtol=RTOL, atol=0) for lam in lams), (res["mw"], lams)
    return res, ref


def test_em_config1_readme_toy():
    sim = simdata.make_config(1)
    D = simdata.host_data(sim)
    init = simdata.init_phi(sim["S"], sim["k"], sim["starts"], seed=1)
    check_em(sim, D, init, complete=False)


@pytest.mark.parametrize("batch", [1, 3])
def test_em_small_gaussian(batch):
    sim = simdata.make_config(dict(S=6, Egenes=5, C=1500, mw=[0.3, 0.3, 0.4], k=3, starts=5, uninform=2), seed=12)
    D = simdata.host_data(sim) + np.random.default_rng(3).normal(size=(sim["E"], sim["C"])) * 0.5
    init = simdata.init_phi(sim["S"], 3, 5, seed=2)
    check_em(sim, D, init, complete=False, batch=batch)


def test_em_many_pairs_few_cells():
    """20 and 30 (start, component) pairs per pass on a small matrix: the E-step then splits the E-gene range over up to
    4 warps of a CTA and combines their partial sums through shared memory.  With k = 5 on 1200 cells some greedy
    decisions are rounding-level ties (oracle margin 1e-16), so the fit is REPLAYED through the oracle (every decision
    equal or within 1e-13), and it must not depend on what shares its passes (bit-identical for batch 1, 4, 6)."""
    sim = simdata.make_config(dict(S=8, Egenes=40, C=1200, mw=[0.2] * 5, k=5, starts=6, uninform=0), seed=31)
    D = simdata.host_data(sim) + np.random.default_rng(7).normal(size=(sim["E"], sim["C"])) * 0.5
    init = simdata.init_phi(8, 5, 6, seed=8)
    with mb.Context(D, sim["pert"]) as cx:
        one = mb.em_run(cx, init, complete=True, batch=1, max_iter_guard=12, tape=True)
        for s in range(6):
            o = orc.replay_start(D, sim["pert"], 8, init[s], one["tape"][s], complete=True, max_iter_guard=12)
            assert not o["diverged"], o["first_hard"]
            assert np.array_equal(one["adj"][s], o["adj"]) and np.array_equal(one["theta"][s], o["theta"])
            assert one["n_iter"][s] == o["n_iter"] and abs(one["ll"][s] - o["ll"]) <= RTOL * abs(o["ll"])
            close(one["probs"][s], o["probs"], rtol=RTOL, atol=1e-9)
        for batch in (4, 6):
            r = mb.em_run(cx, init, complete=True, batch=batch, max_iter_guard=12)
            for key in ("adj", "theta", "ll", "mw", "n_iter", "probs"):
                assert np.array_equal(r[key], one[key]), (batch, key)
        s = 2
        want = orc.get_probs(D, sim["pert"], 8, one["adj"][s], one["theta"][s], one["probs"][s], one["mw"][s], complete=True)
        res = [dict(adj=one["adj"][s][i], subtopo=one["theta"][s][i]) for i in range(5)]
        got = mb.getProbs(one["probs"][s], 5, cx, res, mw=one["mw"][s], complete=True)
        close(got["probs"], want["probs"], rtol=1e-11, atol=1e-11)
        close(one["probs"][s], want["probs"], rtol=1e-11, atol=1e-11)


def test_em_hard_assignments():
    """mnem(affinity = 1): every getAffinity call of the loop returns 0/1 assignments (R/mnems.r:1393-1417) -- the
    M-step weights, the mixture-weight updates and the final lambda."""
    sim = simdata.make_config(dict(S=6, Egenes=5, C=1500, mw=[0.3, 0.3, 0.4], k=3, starts=5, uninform=2), seed=12)
    D = simdata.host_data(sim) + np.random.default_rng(3).normal(size=(sim["E"], sim["C"])) * 0.5
    init = simdata.init_phi(sim["S"], 3, 5, seed=2)
    res, ref = check_em(sim, D, init, complete=False, affinity=1)
    soft = mb.mnem(D, sim["pert"], phi=init)
    assert not np.array_equal(res["n_iter"], soft["n_iter"]) or not np.allclose(res["ll"], soft["ll"])   # a different fit
    with pytest.raises(mb.MnemcuError):
        mb.mnem(D, sim["pert"], phi=init, complete=True, affinity=1)     # an error in the reference too


def test_em_config2_shape_complete():
    """BASELINE config 2 shape (S=10, E=100, ~10k cells, k=3), 6 starts, complete = TRUE."""
    sim = simdata.make_config(2)
    D = simdata.host_data(sim)
    init = simdata.init_phi(sim["S"], 3, 6, seed=4)
    check_em(sim, D, init, complete=True)


def test_em_few_cells_100_inner_iterations():
    """C <= 100 switches getProbs to 100 inner iterations (R/mnems_low.r:584-588)."""
    sim = simdata.make_config(dict(S=3, Egenes=4, C=90, mw=[0.5, 0.5], k=2, starts=3, uninform=0), seed=5)
    D = simdata.host_data(sim)
    assert sim["C"] <= 100
    init = simdata.init_phi(3, 2, 3, seed=6)
    check_em(sim, D, init, complete=False)


def test_mnem_k1_branch():
    sim = simdata.make_config(dict(S=5, Egenes=6, C=800, mw=[1.0], k=1, starts=1, uninform=0), seed=8)
    D = simdata.host_data(sim)
    start = simdata.init_phi(5, 1, 1, seed=1)[0, 0]
    res = mb.mnem(D, sim["pert"], phi=start[None, None])
    R = orc.do_mean(D, sim["pert"], 5)
    fit = orc.nem_greedy(R, start)
    pr = orc.get_probs(D, sim["pert"], 5, fit["adj"][None], fit["theta"][None], np.zeros((1, sim["C"])), np.ones(1))
    assert np.array_equal(res["comp"][0]["phi"], fit["adj"]) and np.array_equal(res["comp"][0]["theta"], fit["theta"])
    assert abs(res["ll"] - pr["ll"]) <= RTOL * abs(pr["ll"])


# ---- data context ------------------------------------------------------------------------------------------------
def test_device_generator_matches_host_generator():
    sim = simdata.make_config(dict(S=7, Egenes=9, C=3000, mw=[0.4, 0.6], k=2, starts=1, uninform=5), seed=33)
    D = simdata.host_data(sim)
    with mb.Context.synthetic(sim) as cx:
        assert np.array_equal(cx.read_data(), D)
        assert np.array_equal(cx.read_data(1234, 77), D[:, 1234:1311])
        R = mb.doMean(cx)
    close(R, orc.do_mean(D, sim["pert"], sim["S"]), rtol=1e-11, atol=1e-11)
    with mb.Context(D, sim["pert"]) as cx:
        assert np.array_equal(cx.read_data(), D)


def test_context_from_device_resident_data():
    """mnemcu_ctx_create_device: D already in device memory (a torch tensor here) gives the same context as the host path."""
    import torch
    sim = simdata.make_config(dict(S=6, Egenes=7, C=2500, mw=[0.5, 0.5], k=2, starts=1, uninform=3), seed=41)
    D = simdata.host_data(sim)
    Dd = torch.from_numpy(np.ascontiguousarray(D.T)).cuda()            # [C, E] row-major == E x C column-major
    with mb.Context.from_device(Dd.data_ptr(), sim["E"], sim["C"], sim["pert"]) as cx, mb.Context(D, sim["pert"]) as cy:
        assert np.array_equal(cx.read_data(), D)
        w = np.random.default_rng(0).random(sim["C"])
        assert np.array_equal(mb.doMean(cx, w), mb.doMean(cy, w))
    with pytest.raises(mb.MnemcuError):
        mb.Context.from_device(D.ctypes.data, sim["E"], sim["C"], sim["pert"])      # a host pointer is refused


def test_ragged_segments_and_missing_padding():
    """Segments of very different sizes (1 cell, 63, 64, 65, 1000 cells) exercise the 64-cell padding."""
    rng = np.random.default_rng(9)
    sizes = [1, 63, 64, 65, 1000, 2]
    pert = np.concatenate([np.full(n, i + 1) for i, n in enumerate(sizes)]).astype(np.int32)
    pert = pert[rng.permutation(len(pert))]
    S, E, k = len(sizes), 37, 2
    D = rng.normal(size=(E, len(pert)))
    res = [dict(adj=rand_graph(rng, S, 0.3), subtopo=rng.integers(0, S + 1, E).astype(np.int32)) for _ in range(k)]
    L_in = np.zeros((k, len(pert)))
    got = mb.getProbs(L_in, k, D, res, pert=pert)
    want = orc.get_probs(D, pert, S, np.stack([r["adj"] for r in res]), np.stack([r["subtopo"] for r in res]), L_in, None)
    close(got["probs"], want["probs"], rtol=1e-11, atol=1e-11)
    assert abs(got["ll"] - want["ll"]) <= RTOL * abs(want["ll"])
    w = rng.random(len(pert))
    close(mb.doMean(D, w, pert), orc.do_mean(D, pert, S, w), rtol=1e-11, atol=1e-11)


def test_bootstrap_matches_oracle_loop():
    """bootstrap() R/mnems.r:2348-2369: size x k searches as one device batch vs one oracle search per resample."""
    sim = simdata.make_config(dict(S=5, Egenes=6, C=900, mw=[0.5, 0.5], k=2, starts=2, uninform=0), seed=14)
    D = simdata.host_data(sim) + np.random.default_rng(1).normal(size=(sim["E"], sim["C"])) * 0.3
    init = simdata.init_phi(5, 2, 2, seed=3)
    fit = mb.mnem(D, sim["pert"], phi=init)
    support, idx = mb.bootstrap(fit, D, sim["pert"], size=12, p=0.8, seed=5)
    post = orc.get_affinity(fit["probs"], fit["mw"])
    for i in range(2):
        R = orc.do_mean(D, sim["pert"], 5, post[i])
        acc = np.zeros((5, 5))
        for j in range(12):
            r = orc.nem_greedy(R[idx[i, j]], None)
            assert r["min_margin"] > 1e-13
            acc += orc.mytc(r["adj"])
        assert np.array_equal(support[i], acc / 12)


# ---- Rho path: cells with several knock-downs ------------------------------------------------
@pytest.mark.parametrize("S,Eg,Cn,seed", [(4, 3, 300, 1), (6, 5, 2000, 2), (12, 8, 5000, 3)])
def test_rho_path_do_mean_and_get_probs(S, Eg, Cn, seed):
    """doMean with normalised Rho columns (R/mnems_low.r:896-904), E-step with OR-ed effect rows (:613-616), theta
    given and theta = NULL, on data with single, double and triple knock-downs."""
    m = simdata.make_multi(S=S, Egenes=Eg, C=Cn, mw=(0.5, 0.5), seed=seed)
    D, pe, E, k = m["D"], m["pert_enc"], m["E"], 2
    rng = np.random.default_rng(seed)
    w = rng.random((3, Cn))
    w[rng.random((3, Cn)) < 0.05] = 0.0
    with mb.Context(D, m["labels"]) as cx:                       # "_" labels -> getRho -> mnemcu_ctx_create_rho
        assert cx.S == S
        R = mb.doMean(cx, w)
        R0 = mb.doMean(cx)
        for j in range(3):
            close(R[j], orc.do_mean(D, pe, S, w[j]), rtol=1e-11, atol=1e-11)
        close(R0, orc.do_mean(D, pe, S), rtol=1e-11, atol=1e-11)
        res = [dict(adj=orc.transitive_reduction(rand_graph(rng, S, 0.3)), subtopo=rng.integers(0, S + 1, E).astype(np.int32))
               for _ in range(k)]
        L_in = rng.normal(size=(k, Cn))
        got = mb.getProbs(L_in, k, cx, res, complete=True)
        want = orc.get_probs(D, pe, S, np.stack([r["adj"] for r in res]), np.stack([r["subtopo"] for r in res]), L_in, None,
                             complete=True)
        close(got["probs"], want["probs"], rtol=1e-11, atol=1e-11)
        assert abs(got["ll"] - want["ll"]) <= RTOL * abs(want["ll"])
        phi = simdata.init_phi(S, k, 1, seed=seed)[0]
        L0 = np.zeros((k, Cn))
        got = mb.getProbs(L0, k, cx, [dict(adj=p) for p in phi])
        want = orc.get_probs(D, pe, S, phi, None, L0, None)
        assert np.array_equal(got["theta_used"], want["theta_used"])
        close(got["probs"], want["probs"], rtol=1e-10, atol=1e-10)
        assert abs(got["ll"] - want["ll"]) <= RTOL * abs(want["ll"])
    # the same context from an explicit Rho matrix
    close(mb.doMean(D, w[0], Rho=m["Rho"]), orc.do_mean(D, pe, S, w[0]), rtol=1e-11, atol=1e-11)


def test_rho_path_with_control_cells_and_empty_sgenes():
    """Cells without any knock-down (all-zero Rho column) have no effects and enter no S-gene's mean; an S-gene that
    is only ever knocked down together with others has no segment of its own."""
    m = simdata.make_multi(S=5, Egenes=4, C=400, seed=9)
    D, Rho = m["D"], m["Rho"].copy()
    Rho[:, :25] = 0.0                                             # control cells
    only_multi = Rho[4] * (Rho.sum(0) == 1) > 0                   # give S-gene 5 a partner wherever it is alone
    Rho[0, only_multi] = 1.0
    masks = (Rho.T @ (1 << np.arange(5))).astype(np.int64)
    single = np.array([bin(int(x)).count("1") == 1 for x in masks])
    pe = np.where(single, np.log2(np.maximum(masks, 1)).astype(np.int64) + 1, -masks).astype(np.int32)
    rng = np.random.default_rng(4)
    w = rng.random(400)
    close(mb.doMean(D, w, Rho=Rho), orc.do_mean(D, pe, 5, w), rtol=1e-11, atol=1e-11)
    res = [dict(adj=rand_graph(rng, 5, 0.3), subtopo=rng.integers(0, 6, m["E"]).astype(np.int32)) for _ in range(2)]
    got = mb.getProbs(np.zeros((2, 400)), 2, D, res, Rho=Rho)
    want = orc.get_probs(D, pe, 5, np.stack([r["adj"] for r in res]), np.stack([r["subtopo"] for r in res]),
                         np.zeros((2, 400)), None)
    close(got["probs"], want["probs"], rtol=1e-11, atol=1e-11)
    assert (got["probs"][:, :25] == 0).all()


@pytest.mark.parametrize("S,Eg,Cn,seed", [(5, 4, 900, 21), (7, 6, 3000, 22)])
def test_rho_path_em(S, Eg, Cn, seed):
    """mnem() on multi-knock-down data: graphs and theta bit-exact, likelihoods to 1e-9 (same bar as single knock-downs)."""
    m = simdata.make_multi(S=S, Egenes=Eg, C=Cn, mw=(0.45, 0.55), seed=seed)
    sim = dict(S=S, pert=m["pert_enc"])
    init = simdata.init_phi(S, 2, 4, seed=seed)
    check_em(sim, m["D"], init, complete=False, gpu_pert=m["labels"])


# ---- marginal = TRUE scoring (R/mnems.r:455-457) ------------------------------------------------
@pytest.mark.parametrize("S,E,logtype", [(3, 10, 2.0), (8, 100, 2.0), (20, 333, np.e), (30, 600, 2.0)])
def test_score_adj_marginal(S, E, logtype):
    rng = np.random.default_rng(S + E)
    R = rng.normal(size=(E, S)) * 3
    adjs = np.stack([rand_graph(rng, S, rng.uniform(0.02, 0.5)) for _ in range(20)])
    for tc in (True, False):
        got = mb.scoreAdj(R, adjs, trans_close=tc, marginal=True, logtype=logtype)
        for i, a in enumerate(adjs):
            sc, th, W = orc.score_adj(R, a, trans_close=tc, marginal=True, logtype=logtype)
            assert np.array_equal(got["subweights"][i], W)
            assert np.array_equal(got["subtopo"][i], th)
            assert abs(got["score"][i] - sc) <= RTOL * abs(sc)


@pytest.mark.parametrize("S,E,n,logtype", [(3, 10, 4, 2.0), (6, 40, 6, 2.0), (10, 100, 4, np.e), (16, 130, 2, 2.0)])
def test_nem_greedy_marginal(S, E, n, logtype):
    """nem(marginal = TRUE): same moves, same graph, same theta as the oracle; scores to 1e-9 (log / pow are libm on one
    side and CUDA math on the other, so a decision must be clearer than 1e-9 for the case to be admissible)."""
    rng = np.random.default_rng(2000 + S)
    R = np.stack([planted_R(rng, S, E, snr=1.0) / 4 for _ in range(n)])      # small values: b^W stays finite
    starts = np.stack([orc.transitive_reduction(rand_graph(rng, S, 0.15, dag=True)) if i % 2 else np.zeros((S, S))
                       for i in range(n)])
    got = mb.nem_greedy(R, starts, marginal=True, logtype=logtype)
    for i in range(n):
        want = orc.nem_greedy(R[i], starts[i], marginal=True, logtype=logtype)
        assert want["min_margin"] > 1e-9, "near-tie between candidates; pick another seed"
        assert np.array_equal(got["adj"][i], want["adj"])
        assert np.array_equal(got["subtopo"][i], want["theta"])
        assert got["n_iter"][i] == want["n_iter"] and got["n_scored"][i] == want["n_raw"]
        assert abs(got["score"][i] - want["score"]) <= RTOL * abs(want["score"])


# ---- the boundary from a compiled C++ program -------------------------------------------------------------------------
def test_cpp_consumer_fit_matches_the_python_binding(tmp_path):
    """tests/cabi/cabi_check.cpp (plain g++, links libmnemcu.so) runs mnemcu_em_run on column-major buffers read from a
    file, as the Rcpp shim would pass them; winner, likelihood, iteration count and graphs equal the ctypes path."""
    import subprocess
    exe = os.path.join(os.path.dirname(__file__), "cabi", "cabi_check")
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build()
    sim = simdata.make_config(dict(S=5, Egenes=6, C=1500, mw=[0.4, 0.6], k=2, starts=3, uninform=1), seed=51)
    D = simdata.host_data(sim)
    init = simdata.init_phi(5, 2, 3, seed=6)
    path = str(tmp_path / "problem.bin")
    with open(path, "wb") as f:
        f.write(np.array([sim["E"], sim["C"], 5, 2, 3, 0], dtype=np.int32).tobytes())
        f.write(np.asfortranarray(D).tobytes(order="F"))
        f.write(np.ascontiguousarray(sim["pert"], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(np.transpose(init.reshape(6, 5, 5), (0, 2, 1))).astype(np.uint8).tobytes())
    r = subprocess.run([exe, "fit", path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    tok = r.stdout.split()
    best, ll, iters, digest = int(tok[1]), float(tok[3]), int(tok[5]), int(tok[7])
    fit = mb.mnem(D, sim["pert"], phi=init)
    assert best == fit["best_start"] and ll == fit["ll"] and iters == int(fit["n_iter"].sum())
    dig = 1469598103934665603
    adj = np.ascontiguousarray(np.transpose(fit["limits"][best]["adj"], (0, 2, 1))).astype(np.uint8).ravel()
    th = fit["limits"][best]["theta"].astype(np.uint32).ravel()
    for v in list(adj) + list(th):
        dig = ((dig ^ int(v)) * 1099511628211) % (1 << 64)
    assert dig == digest


def test_rcpp_shim_fit_matches_the_python_binding(tmp_path):
    """rshim/mnemcu_shim.cpp -- the Rcpp glue of INTEGRATION.md -- compiled against the functional Rcpp stand-in
    (tests/cabi/rcpp_stub/Rcpp.h) and driven like rshim/mnem_cuda.R drives it: mnemcu_context + mnemcu_em on mnem()'s
    `phi` list.  Winner, likelihood, iteration count and graphs equal the ctypes path; the legacy transClose_W symbol and
    the external-pointer finalizer work."""
    import subprocess
    exe = os.path.join(os.path.dirname(__file__), "cabi", "shim_check")
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build()
    sim = simdata.make_config(dict(S=5, Egenes=6, C=1500, mw=[0.4, 0.6], k=2, starts=3, uninform=1), seed=51)
    D = simdata.host_data(sim)
    init = simdata.init_phi(5, 2, 3, seed=6)
    path = str(tmp_path / "problem.bin")
    with open(path, "wb") as f:
        f.write(np.array([sim["E"], sim["C"], 5, 2, 3, 0], dtype=np.int32).tobytes())
        f.write(np.asfortranarray(D).tobytes(order="F"))
        f.write(np.ascontiguousarray(sim["pert"], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(np.transpose(init.reshape(6, 5, 5), (0, 2, 1))).astype(np.uint8).tobytes())
    r = subprocess.run([exe, "fit", path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    tok = lines[0].split()
    best, ll, iters, digest = int(tok[1]), float(tok[3]), int(tok[5]), int(tok[7])
    fit = mb.mnem(D, sim["pert"], phi=init)
    assert best == fit["best_start"] and ll == fit["ll"] and iters == int(fit["n_iter"].sum())
    dig = 1469598103934665603
    adj = np.ascontiguousarray(np.transpose(fit["limits"][best]["adj"], (0, 2, 1))).astype(np.uint8).ravel()
    th = fit["limits"][best]["theta"].astype(np.uint32).ravel()
    for v in list(adj) + list(th):
        dig = ((dig ^ int(v)) * 1099511628211) % (1 << 64)
    assert dig == digest
    assert lines[1] == "closure_1_3 1" and lines[2] == "ctx_released 1"


def test_user_mixture_weights_and_winner_probs():
    """mnem(mw =) (R/mnems.r:1771-1773: the initial mixture weights of every start) replayed through the oracle, and
    probs = 'winner' (one k x C block, the best start's) against the per-start layout."""
    sim = simdata.make_config(dict(S=6, Egenes=8, C=3000, mw=[0.3, 0.7], k=2, starts=4, uninform=2), seed=77)
    D = simdata.host_data(sim)
    init = simdata.init_phi(6, 2, 4, seed=9)
    mw0 = np.array([0.85, 0.15])
    with mb.Context(D, sim["pert"]) as cx:
        r = mb.em_run(cx, init, mw=mw0, tape=True)
        rw = mb.em_run(cx, init, mw=mw0, probs="winner")
        r0 = mb.em_run(cx, init)
    for s in range(4):
        o = orc.replay_start(D, sim["pert"], 6, init[s], r["tape"][s], mw=mw0)
        assert not o["diverged"], o["first_hard"]
        assert np.array_equal(r["adj"][s], o["adj"]) and np.array_equal(r["theta"][s], o["theta"])
        assert r["n_iter"][s] == o["n_iter"]
        assert abs(r["ll"][s] - o["ll"]) <= 1e-9 * abs(o["ll"])
        assert np.allclose(r["mw"][s], o["mw"], rtol=1e-9, atol=0)
    assert r0["lls"][0][0] != r["lls"][0][0]                           # the weights do change the path of the fit
    assert rw["best"] == r["best"] and rw["probs"].shape[0] == 1
    assert np.array_equal(rw["probs"][0], r["probs"][r["best"]])
    assert np.array_equal(rw["adj"], r["adj"]) and np.array_equal(rw["ll"], r["ll"])      # the tape does not change a fit


def test_dictionary_overflow_falls_back_to_the_generic_scorer(monkeypatch):
    """More distinct column masks in a neighbourhood than the shared-memory table holds must not fail the call: the
    searches are run again with the generic scorer (every changed column re-summed, no dictionary).  Forced here by
    capping the table at 8 entries; graphs, theta and iteration counts equal the normal path and the oracle."""
    rng = np.random.default_rng(12)
    S, E, n = 9, 150, 5
    R = rng.normal(size=(n, E, S)) * 3
    normal = mb.nem_greedy(R, None)
    monkeypatch.setenv("MNEMCU_DICT_CAP", "8")
    capped = mb.nem_greedy(R, None)
    monkeypatch.delenv("MNEMCU_DICT_CAP")
    for q in range(n):
        ref = orc.nem_greedy(R[q], None)
        assert ref["min_margin"] > 1e-13
        assert np.array_equal(capped["adj"][q], ref["adj"]) and np.array_equal(capped["subtopo"][q], ref["theta"])
        assert np.array_equal(capped["adj"][q], normal["adj"][q]) and capped["n_iter"][q] == normal["n_iter"][q]
        assert abs(capped["score"][q] - ref["score"]) <= 1e-9 * abs(ref["score"])


def test_em_with_many_components():
    """k up to MNEMCU_MAX_K = 32 components (one start then fills a whole pass): k = 24 replayed through the oracle."""
    sim = simdata.make_config(dict(S=5, Egenes=4, C=2400, mw=[1 / 24] * 24, k=24, starts=2, uninform=0), seed=31)
    D = simdata.host_data(sim)
    init = simdata.init_phi(5, 24, 2, seed=8)
    with mb.Context(D, sim["pert"]) as cx:
        r = mb.em_run(cx, init, tape=True)
    for s in range(2):
        o = orc.replay_start(D, sim["pert"], 5, init[s], r["tape"][s])
        assert not o["diverged"], o["first_hard"]
        assert np.array_equal(r["adj"][s], o["adj"]) and np.array_equal(r["theta"][s], o["theta"])
        assert abs(r["ll"][s] - o["ll"]) <= 1e-9 * abs(o["ll"])


def test_null_component():
    """mnem(nullcomp = TRUE) (R/mnems.r:1832-1834, 1894-1900, 2012, 2076-2080; R/mnems_low.r:607-608): component k + 1 has
    an empty graph, theta 0 and log ratios 0, and takes part in the mixture only.  Replayed through the oracle."""
    sim = simdata.make_config(dict(S=6, Egenes=8, C=3000, mw=[0.5, 0.5], k=2, starts=3, uninform=2), seed=41)
    D = simdata.host_data(sim)
    D[:, ::7] = np.random.default_rng(3).normal(size=D[:, ::7].shape)          # cells no component explains: the null one takes them
    init2 = simdata.init_phi(6, 2, 3, seed=5)
    fit = mb.mnem(D, sim["pert"], phi=init2, nullcomp=True)
    assert len(fit["comp"]) == 3 and not fit["comp"][2]["phi"].any() and not fit["comp"][2]["theta"].any()
    assert np.all(fit["probs"][2] == 0.0) and fit["mw"][2] > 1e-4
    init3 = np.concatenate([init2, np.zeros_like(init2[:, :1])], axis=1)
    with mb.Context(D, sim["pert"]) as cx:
        r = mb.em_run(cx, init3, nullcomp=True, tape=True)
    for s in range(3):
        o = orc.replay_start(D, sim["pert"], 6, init3[s], r["tape"][s], nullcomp=True)
        assert not o["diverged"], o["first_hard"]
        assert np.array_equal(r["adj"][s], o["adj"]) and np.array_equal(r["theta"][s], o["theta"])
        assert r["n_iter"][s] == o["n_iter"] and abs(r["ll"][s] - o["ll"]) <= RTOL * abs(o["ll"])
        close(r["mw"][s], o["mw"], rtol=RTOL)
        close(r["probs"][s], o["probs"], rtol=RTOL, atol=1e-9)
    assert r["best"] == fit["best_start"] and np.array_equal(r["adj"][r["best"]][2], np.zeros((6, 6)))


def test_fixed_theta():
    """mnem(theta =) (R/mnems.r:1890-1892, 2007-2011): the E-gene attachments are given and never move -- the first
    getProbs takes its theta-given path, every M-step still searches the graphs (the llr score ignores subtopo,
    R/mnems.r:439-454) and returns the given theta.  Replayed through the oracle."""
    sim = simdata.make_config(dict(S=6, Egenes=8, C=2500, mw=[0.4, 0.6], k=2, starts=3, uninform=2), seed=23)
    D = simdata.host_data(sim)
    init = simdata.init_phi(6, 2, 3, seed=4)
    rng = np.random.default_rng(2)
    th = np.stack([np.where(rng.random(sim["theta_true"].shape) < 0.8, sim["theta_true"], rng.integers(0, 7, sim["theta_true"].shape))
                   for _ in range(3)]).astype(np.int32)
    with mb.Context(D, sim["pert"]) as cx:
        r = mb.em_run(cx, init, theta=th, tape=True)
        free = mb.em_run(cx, init)
    assert np.array_equal(r["theta"], th) and all((t == 0).all() for t in r["thetaevo"])
    assert not np.array_equal(free["theta"], th)
    for s in range(3):
        o = orc.replay_start(D, sim["pert"], 6, init[s], r["tape"][s], theta=th[s])
        assert not o["diverged"], o["first_hard"]
        assert np.array_equal(r["adj"][s], o["adj"]) and np.array_equal(o["theta"], th[s])
        assert r["n_iter"][s] == o["n_iter"] and abs(r["ll"][s] - o["ll"]) <= RTOL * abs(o["ll"])
        assert np.array_equal(r["phievo"][s], o["phievo"])
        close(r["lls"][s], o["lls"], rtol=RTOL)
        close(r["mw"][s], o["mw"], rtol=RTOL)
        close(r["probs"][s], o["probs"], rtol=RTOL, atol=1e-9)
    ref = orc.mnem_em(D, sim["pert"], 6, init, theta=th)
    assert ref["min_margin"].min() <= 1e-13 or all(np.array_equal(r["adj"][s], ref["adj"][s]) for s in range(3))


@pytest.mark.parametrize("S,Egenes,Cn,k,uninform", [(2, 1, 7, 2, 0), (2, 1, 64, 2, 1), (3, 2, 65, 2, 0), (3, 1, 130, 3, 2), (4, 3, 33, 2, 0),
                                                   (5, 1, 200, 4, 0), (7, 2, 129, 2, 3)])
def test_tiny_and_ragged_shapes(S, Egenes, Cn, k, uninform):
    """Degenerate sizes: a single E-gene per S-gene, fewer cells than a 64-cell block, cell counts one past a block, more
    components than the data support.  Every fit is replayed through the oracle (C <= 100 also switches getProbs to its
    100 inner iterations, R/mnems_low.r:584-588)."""
    sim = simdata.make_config(dict(S=S, Egenes=Egenes, C=Cn, mw=[1.0 / k] * k, k=k, starts=3, uninform=uninform), seed=100 + S + Cn)
    D = simdata.host_data(sim)
    init = simdata.init_phi(S, k, 3, seed=S * Cn)
    with mb.Context(D, sim["pert"]) as cx:
        r = mb.em_run(cx, init, tape=True, max_iter_guard=30)
        rb = mb.em_run(cx, init, batch=1, max_iter_guard=30)
    assert np.array_equal(r["adj"], rb["adj"]) and np.array_equal(r["ll"], rb["ll"])
    for s in range(3):
        o = orc.replay_start(D, sim["pert"], S, init[s], r["tape"][s], max_iter_guard=30)
        assert not o["diverged"], o["first_hard"]
        assert np.array_equal(r["adj"][s], o["adj"]) and np.array_equal(r["theta"][s], o["theta"])
        assert r["n_iter"][s] == o["n_iter"]
        if np.isfinite(o["ll"]):
            assert abs(r["ll"][s] - o["ll"]) <= RTOL * max(abs(o["ll"]), 1e-300)
            close(r["probs"][s], o["probs"], rtol=RTOL, atol=1e-9)
        else:
            assert r["ll"][s] == o["ll"] or (np.isnan(r["ll"][s]) and np.isnan(o["ll"]))
