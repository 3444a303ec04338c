"""Cross-checks the C oracle against an independent, vectorised NumPy reading of the same reference formulas
(SURVEY.md appendix A), on small random cases including the edge cases the reference code handles:
non-closed intermediate graphs, cycles, unattached E-genes, k = 1, complete = TRUE."""
import numpy as np
import pytest

import oracle as orc
from mnem_b200 import simdata


def np_closure(a):
    a = (np.asarray(a) != 0).astype(int)
    np.fill_diagonal(a, 1)
    n = len(a)
    for _ in range(n):
        a = ((a + a @ a) > 0).astype(int)
    return a


def np_reduction(g):
    """R/mnems.r:517-532 transcribed literally (order-dependent)."""
    g = np_closure(g)
    np.fill_diagonal(g, 0)
    n = len(g)
    for y in range(n):
        for x in range(n):
            if g[x, y] != 0:
                for j in range(n):
                    if g[y, j] != 0:
                        g[x, j] = 0
    return g


def rand_graph(rng, S, p=0.25, dag=False):
    a = (rng.random((S, S)) < p).astype(np.uint8)
    if dag:
        a = np.triu(a, 1)
    np.fill_diagonal(a, 0)
    return a


@pytest.mark.parametrize("S", [2, 3, 5, 8, 13])
def test_closure_and_reduction(S):
    rng = np.random.default_rng(S)
    for _ in range(30):
        g = rand_graph(rng, S, p=rng.uniform(0.05, 0.5), dag=rng.random() < 0.5)
        assert (orc.mytc(g) == np_closure(g)).all()
        assert (orc.transitive_reduction(g) == np_reduction(g)).all()
        u, v = rng.integers(1, S + 1, 2)
        x = np_closure(g)
        want = x | np.outer(x[:, u - 1], x[v - 1, :])
        assert (orc.trans_close_ins(x, u, v) == want).all()


def np_neighbours(P):
    """R/mnems_low.r:1025-1084 transcribed with dense matrices; returns the list before unique()."""
    P = np.array(P, dtype=int)
    S = len(P)
    out = []
    for idx in range(S * S):                                  # which(Phi == 0), column-major
        u, v = idx % S, idx // S
        if P[u, v] == 0:
            m = P.copy()
            m[u, v] = 1
            np.fill_diagonal(m, 1)
            m = m | np.outer(m[:, u], m[v, :])
            out.append(m)
    T = np_reduction(P)
    for c in range(S):
        for r in range(S):
            if T[r, c] + T[c, r] == 1:
                m = P.copy()
                m[r, c] = 1 - m[r, c]
                m[c, r] = 1 - m[c, r]
                np.fill_diagonal(m, 1)
                u, v = (r, c) if m[r, c] == 1 else (c, r)
                if m[u, v] == 1:
                    m = m | np.outer(m[:, u], m[v, :])
                else:
                    m[u, v] = int((m[u, :] & m[:, v]).any())
                out.append(m)
    P0 = P - np.eye(S, dtype=int)
    T = np_reduction(P0)
    for idx in range(S * S):
        u, v = idx % S, idx // S
        if T[u, v] == 1:
            m = P0.copy()
            m[u, v] = 0
            np.fill_diagonal(m, 1)
            out.append(m)
    return out


@pytest.mark.parametrize("S", [2, 4, 6, 9])
def test_neighbours(S):
    rng = np.random.default_rng(100 + S)
    for t in range(12):
        P = rand_graph(rng, S, p=rng.uniform(0.05, 0.45), dag=t % 2 == 0)
        if t % 3 == 0:
            P = np_closure(P)
        np.fill_diagonal(P, 1)
        raw = np_neighbours(P)
        uniq = []
        for m in raw:
            if not any((m == q).all() for q in uniq):
                uniq.append(m)
        got, kind, nraw = orc.neighbours(P)
        assert nraw == len(raw)
        assert len(got) == len(uniq)
        for a, b in zip(got, uniq):
            assert (a == b).all()


def np_estep(D, pert, phi, theta):
    Phi = np_closure(phi)
    E = D.shape[0]
    F = np.zeros((E, Phi.shape[0]))
    att = theta > 0
    F[att, :] = Phi[:, theta[att] - 1].T                      # F[e, s] = Phi[s, theta(e)]
    return np.einsum("ec,ec->c", D, F[:, pert - 1])


def test_estep_theta_given_and_ll():
    rng = np.random.default_rng(5)
    S, E, Cn, k = 6, 17, 143, 3
    D = rng.normal(size=(E, Cn))
    pert = rng.integers(1, S + 1, Cn).astype(np.int32)
    phi = np.stack([rand_graph(rng, S, 0.3) for _ in range(k)])
    theta = rng.integers(0, S + 1, (k, E)).astype(np.int32)
    mw = np.array([0.2, 0.5, 0.3])
    for complete in (False, True):
        r = orc.get_probs(D, pert, S, phi, theta, np.zeros((k, Cn)), mw, complete=complete)
        want = np.stack([np_estep(D, pert, phi[i], theta[i]) for i in range(k)])
        np.testing.assert_allclose(r["probs"], want, rtol=1e-13, atol=1e-13)
        # inner loop: mw fixed point, 10 iterations, ll of the LAST iteration with the mw before its update
        m = orc.mw_update(np.zeros((k, Cn)), mw, complete=complete, na_guard=True)
        ll = None
        for _ in range(10):
            ll = orc.get_ll(want, m, complete=complete)
            m = orc.mw_update(want, m, complete=complete)
        assert abs(r["ll"] - ll) <= 1e-12 * abs(ll)
        np.testing.assert_allclose(r["mw"], m, rtol=1e-12)
        # getAffinity / getLL against plain NumPy
        g = orc.get_affinity(want, mw, complete=complete)
        x = want.copy()
        if complete and (x > 0).any():
            x = x - (x.max(0) - (np.log(2.0 ** 1023) / np.log(2)) / k)
        y = 2.0 ** x * mw[:, None]
        np.testing.assert_allclose(g, y / y.sum(0), rtol=1e-13)
        if complete:
            llw = (g * (want + np.log(mw)[:, None] / np.log(2))).sum()
        else:
            llw = np.log2((2.0 ** want * mw[:, None]).sum(0)).sum()
        assert abs(orc.get_ll(want, mw, complete=complete) - llw) <= 1e-12 * abs(llw)


def test_estep_theta_free_matches_definition():
    rng = np.random.default_rng(6)
    S, E, Cn, k = 5, 11, 140, 2
    D = rng.normal(size=(E, Cn))
    pert = rng.integers(1, S + 1, Cn).astype(np.int32)
    phi = np.stack([rand_graph(rng, S, 0.3, dag=True) for _ in range(k)])
    L0 = np.zeros((k, Cn))
    r = orc.get_probs(D, pert, S, phi, None, L0, None)
    # replay: theta_i = first argmax_j [(D * gamma_i) %*% cbind(Phi[pert,], 0)], null column last
    mw = orc.mw_update(L0, np.ones(k) / k, na_guard=True)
    probs, best, llold = L0, L0, -np.inf
    for _ in range(10):
        post = orc.get_affinity(probs, mw)
        new = np.zeros_like(L0)
        for i in range(k):
            Phi = np_closure(phi[i])
            A = np.concatenate([Phi[pert - 1, :], np.zeros((Cn, 1))], 1)
            W = (D * post[i]) @ A
            th = W.argmax(1) + 1
            th[th == S + 1] = 0
            new[i] = np_estep(D, pert, phi[i], th)
        ll0 = orc.get_ll(new, mw)
        if ll0 > llold:
            best = new
        probs, llold = new, ll0
        mw = orc.mw_update(probs, mw)
    np.testing.assert_allclose(r["probs"], best, rtol=1e-12, atol=1e-12)
    assert abs(r["ll"] - llold) <= 1e-12 * abs(llold)


def test_do_mean_and_score():
    rng = np.random.default_rng(7)
    S, E, Cn = 7, 23, 300
    D = rng.normal(size=(E, Cn))
    pert = rng.integers(1, S + 1, Cn).astype(np.int32)
    w = rng.random(Cn)
    w[rng.random(Cn) < 0.1] = 0.0
    R = orc.do_mean(D, pert, S, w)
    Rho = (pert[None, :] == np.arange(1, S + 1)[:, None]).astype(float)
    np.testing.assert_allclose(R, (D * w) @ Rho.T, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(orc.do_mean(D, pert, S), D @ Rho.T, rtol=1e-12, atol=1e-12)
    for _ in range(10):
        adj = rand_graph(rng, S, 0.3)
        for tc in (True, False):
            sc, th, W = orc.score_adj(R, adj, trans_close=tc)
            A = np_closure(adj) if tc else (adj != 0).astype(int)
            Ww = R @ A
            np.testing.assert_allclose(W, Ww, rtol=1e-12, atol=1e-12)
            full = np.concatenate([np.zeros((E, 1)), W], 1)
            assert (th == full.argmax(1)).all()
            assert abs(sc - full.max(1).sum()) <= 1e-12 * abs(sc) + 1e-12


def np_greedy(R, start):
    """nem(search='greedy') with dense NumPy, full re-scoring of every candidate."""
    S = R.shape[1]
    better = np.array(start, dtype=int)
    np.fill_diagonal(better, 1)

    def score(A):
        return np.maximum((R @ A).max(1), 0).sum()
    old = score(np_closure(better))
    n_it = 0
    while True:
        cands = np_neighbours(better)
        sc = np.array([score(m) for m in cands])
        b = int(sc.argmax())
        if sc[b] > old or (sc[b] == old and better.sum() > cands[b].sum()):
            better, old = cands[b], sc[b]
            n_it += 1
        else:
            break
    A = np_closure(better)
    th = np.concatenate([np.zeros((len(R), 1)), R @ A], 1).argmax(1)
    return np_reduction(better), th, old, n_it


@pytest.mark.parametrize("seed", range(6))
def test_greedy_search(seed):
    rng = np.random.default_rng(40 + seed)
    S, E = int(rng.integers(3, 8)), int(rng.integers(8, 30))
    truth = np_closure(rand_graph(rng, S, 0.3, dag=True))
    theta = rng.integers(1, S + 1, E)
    R = np.where(truth[:, theta - 1].T > 0, 1.0, -1.0) * rng.uniform(5, 20) + rng.normal(size=(E, S)) * 4
    for start in (None, np_reduction(rand_graph(rng, S, 0.2, dag=True))):
        got = orc.nem_greedy(R, start)
        adj, th, sc, n_it = np_greedy(R, np.zeros((S, S)) if start is None else start)
        assert (got["adj"] == adj).all()
        assert (got["theta"] == th).all()
        assert got["n_iter"] == n_it
        assert abs(got["score"] - sc) <= 1e-10 * abs(sc)


def test_em_oracle_properties():
    sim = simdata.make_config(dict(S=4, Egenes=3, C=240, mw=[0.5, 0.5], k=2, starts=4, uninform=1), seed=3)
    D = simdata.host_data(sim)
    init = simdata.init_phi(sim["S"], 2, 4, seed=9)
    r1 = orc.mnem_em(D, sim["pert"], sim["S"], init, nthreads=1)
    r4 = orc.mnem_em(D, sim["pert"], sim["S"], init, nthreads=4, faithful_cost=True)
    for key in ("adj", "theta", "probs", "mw", "ll", "n_iter"):
        assert np.array_equal(r1[key], r4[key]), key          # thread count and redundant recomputation change nothing
    for s in range(4):
        lls = r1["lls"][s]
        assert len(lls) == r1["n_iter"][s] >= 1
        assert (np.diff(lls) >= 0).all()                      # accepted likelihood never decreases (R/mnems.r:2104-2108)
        assert r1["ll"][s] == lls.max()
        np.testing.assert_allclose(r1["mw"][s].sum(), 1.0, rtol=1e-12)
    assert r1["best"] == max(i for i in range(4) if r1["ll"][i] == r1["ll"].max())   # `<=`: last maximal start
    rf = orc.mnem_em(D, sim["pert"], sim["S"], init, acc="f64")
    assert np.array_equal(rf["adj"], r1["adj"]) and np.array_equal(rf["theta"], r1["theta"])
    np.testing.assert_allclose(rf["ll"], r1["ll"], rtol=1e-12)


# ---- Rho path: cells with several knock-downs (R/mnems_low.r:308-324, 613-616, 896-904) ------------------------------
def test_rho_path_matches_the_matrix_formulas():
    """The oracle's multi-knock-down branches against the reference's own matrix expressions evaluated in NumPy:
    doMean = (D * w) %*% t(Rho') with Rho' columns normalised to sum 1; E-step effect rows min(1, t(Rho) %*% adj1);
    theta-free E-step argmax over (D * gamma) %*% cbind(adj1, 0)."""
    m = simdata.make_multi(S=5, Egenes=3, C=260, seed=12)
    D, Rho, pe, S, E, Cn, k = m["D"], m["Rho"], m["pert_enc"], m["S"], m["E"], m["C"], m["k"]
    assert (Rho.sum(0) > 1).any() and (Rho.sum(0) == 1).any()
    rng = np.random.default_rng(2)
    w = rng.random(Cn)
    w[rng.random(Cn) < 0.1] = 0.0
    Rn = Rho / np.maximum(Rho.sum(0), 1)[None, :]
    np.testing.assert_allclose(orc.do_mean(D, pe, S, w), (D * w) @ Rn.T, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(orc.do_mean(D, pe, S), D @ Rn.T, rtol=1e-12, atol=1e-12)
    phi = np.stack([rand_graph(rng, S, 0.3, dag=True) for _ in range(k)])
    theta = rng.integers(0, S + 1, (k, E)).astype(np.int32)

    def estep(phi_i, th):
        adj1 = np.minimum(1, Rho.T @ np_closure(phi_i))                       # C x S
        adj1 = np.concatenate([adj1, np.zeros((Cn, 1))], 1)
        sub = np.where(th == 0, S + 1, th)
        return (D * adj1[:, sub - 1].T).sum(0)                                # colSums(data * t(adj2))

    r = orc.get_probs(D, pe, S, phi, theta, np.zeros((k, Cn)), None)
    want = np.stack([estep(phi[i], theta[i]) for i in range(k)])
    np.testing.assert_allclose(r["probs"], want, rtol=1e-12, atol=1e-12)
    # theta = NULL: one replay of the inner loop
    r0 = orc.get_probs(D, pe, S, phi, None, np.zeros((k, Cn)), None)
    mw = orc.mw_update(np.zeros((k, Cn)), np.ones(k) / k, na_guard=True)
    probs, best, llold = np.zeros((k, Cn)), np.zeros((k, Cn)), -np.inf
    for _ in range(10):
        post = orc.get_affinity(probs, mw)
        new = np.zeros((k, Cn))
        for i in range(k):
            adj1 = np.minimum(1, Rho.T @ np_closure(phi[i]))
            W = (D * post[i]) @ np.concatenate([adj1, np.zeros((Cn, 1))], 1)
            th = W.argmax(1) + 1
            th[th == S + 1] = 0
            new[i] = estep(phi[i], th)
        ll0 = orc.get_ll(new, mw)
        if ll0 > llold:
            best = new
        probs, llold = new, ll0
        mw = orc.mw_update(probs, mw)
    np.testing.assert_allclose(r0["probs"], best, rtol=1e-12, atol=1e-12)
    assert abs(r0["ll"] - llold) <= 1e-12 * abs(llold)


def test_rho_path_reduces_to_the_single_knockdown_path():
    """Encoding single knock-downs as masks must not change a bit of the EM fit."""
    sim = simdata.make_config(dict(S=4, Egenes=3, C=200, mw=[0.5, 0.5], k=2, starts=2, uninform=0), seed=5)
    D = simdata.host_data(sim)
    init = simdata.init_phi(4, 2, 2, seed=1)
    a = orc.mnem_em(D, sim["pert"], 4, init)
    enc = -(1 << (sim["pert"].astype(np.int64) - 1))
    b = orc.mnem_em(D, enc.astype(np.int32), 4, init)
    for key in ("adj", "theta", "probs", "ll", "mw"):
        assert np.array_equal(a[key], b[key]), key


# ---- marginal = TRUE scoring (R/mnems.r:455-457) -------------------------------------------------------------------
@pytest.mark.parametrize("logtype", [2.0, np.e, 10.0])
def test_marginal_score_matches_the_formula(logtype):
    rng = np.random.default_rng(11)
    S, E = 7, 31
    R = rng.normal(size=(E, S)) * 2
    for _ in range(6):
        adj = rand_graph(rng, S, 0.3)
        for tc in (True, False):
            sc, th, W = orc.score_adj(R, adj, trans_close=tc, marginal=True, logtype=logtype)
            A = np_closure(adj) if tc else (adj != 0).astype(int)
            full = np.concatenate([np.zeros((E, 1)), R @ A], 1)
            want = (np.log((logtype ** full).sum(1)) / np.log(logtype)).sum()
            assert abs(sc - want) <= 1e-12 * abs(want)
            assert (th == full.argmax(1)).all()                      # dotopo: same attachment as the max variant
            np.testing.assert_allclose(W, R @ A, rtol=1e-12, atol=1e-12)
    # the marginal search only swaps the score: from the same start it must never end below its starting score
    r = orc.nem_greedy(R, None, marginal=True, logtype=logtype)
    s0 = orc.score_adj(R, np.zeros((S, S)), marginal=True, logtype=logtype)[0]
    assert r["score"] >= s0 and r["max_trace"] == r["score"]


# ---- affinity = 1: hard assignments (R/mnems.r:1393-1417) -----------------------------------------------------------
def test_hard_assignments_in_the_oracle_em():
    rng = np.random.default_rng(21)
    k, Cn = 3, 500
    L = rng.normal(size=(k, Cn)) * 5
    mw = np.array([0.2, 0.5, 0.3])
    post = orc.get_affinity(L, mw, affinity=1)
    y = 2.0 ** L * mw[:, None]
    assert np.array_equal(post, (y == y.max(0, keepdims=True)).astype(float))      # one 1 per cell (ties: several)
    lam = orc.final_mw(L, mw, affinity=1)
    np.testing.assert_allclose(lam, post.sum(1) / post.sum(), rtol=1e-15)
    # the EM loop with hard assignments runs, differs from the soft fit, and keeps its invariants
    sim = simdata.make_config(dict(S=4, Egenes=4, C=400, mw=[0.5, 0.5], k=2, starts=3, uninform=0), seed=9)
    D = simdata.host_data(sim) + rng.normal(size=(sim["E"], sim["C"])) * 0.5
    init = simdata.init_phi(4, 2, 3, seed=3)
    hard = orc.mnem_em(D, sim["pert"], 4, init, affinity=1)
    soft = orc.mnem_em(D, sim["pert"], 4, init, affinity=0)
    assert not np.allclose(hard["ll"], soft["ll"])
    for s in range(3):
        assert (np.diff(hard["lls"][s]) >= 0).all()
        np.testing.assert_allclose(hard["mw"][s].sum(), 1.0, rtol=1e-12)
