"""Hypothesis-driven properties of the CPU oracle on small random cases (S <= 7, E <= 14): the algebra the reference's
graph routines must satisfy whatever the input, independent of the NumPy re-reading in test_oracle_selfcheck.py."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle as orc


@st.composite
def graphs(draw, dag=False, max_s=7):
    S = draw(st.integers(1, max_s))
    bits = draw(st.lists(st.booleans(), min_size=S * S, max_size=S * S))
    a = np.array(bits, dtype=np.uint8).reshape(S, S)
    if dag:
        a = np.triu(a, 1)
        p = np.array(draw(st.permutations(list(range(S)))))
        a = a[np.ix_(p, p)]
    np.fill_diagonal(a, 0)
    return a


@settings(max_examples=120, deadline=None)
@given(graphs())
def test_closure_is_reflexive_transitive_and_idempotent(g):
    c = orc.mytc(g)
    S = len(g)
    assert (np.diag(c) == 1).all() and (c >= (g != 0)).all()
    assert ((c.astype(int) @ c.astype(int) > 0) <= (c > 0)).all()            # transitive
    assert np.array_equal(orc.mytc(c), c)
    reach = np.linalg.matrix_power((g != 0).astype(np.int64) + np.eye(S, dtype=np.int64), S) > 0
    assert np.array_equal(c > 0, reach)


@settings(max_examples=120, deadline=None)
@given(graphs(dag=True))
def test_reduction_of_a_dag_is_minimal_and_keeps_the_closure(g):
    r = orc.transitive_reduction(g)
    assert (np.diag(r) == 0).all()
    assert np.array_equal(orc.mytc(r), orc.mytc(g))                            # same reachability
    for i, j in zip(*np.nonzero(r)):                                           # no edge is implied by the others
        q = r.copy()
        q[i, j] = 0
        assert orc.mytc(q)[i, j] == 0
    assert np.array_equal(orc.transitive_reduction(orc.mytc(g)), r)


@settings(max_examples=80, deadline=None)
@given(graphs(max_s=6), st.data())
def test_one_step_insert_closure(g, data):
    """transClose_Ins(u, v) after setting x[u, v]: x[i, j] |= x[i, u] & x[v, j] (src/mm.cpp:50-65); on a closed graph the
    result is closed again."""
    S = len(g)
    x = orc.mytc(g)
    u = data.draw(st.integers(1, S))
    v = data.draw(st.integers(1, S))
    y = x.copy()
    y[u - 1, v - 1] = 1
    got = orc.trans_close_ins(y, u, v)
    want = y | np.outer(y[:, u - 1], y[v - 1, :])
    assert np.array_equal(got, want)
    assert np.array_equal(orc.mytc(got), got)


@settings(max_examples=60, deadline=None)
@given(graphs(max_s=6), st.integers(0, 2 ** 31 - 1))
def test_neighbours_are_unique_ordered_and_single_edge_moves(g, seed):
    P = (g != 0).astype(np.uint8)
    if seed % 2:
        P = orc.mytc(P)
    np.fill_diagonal(P, 1)
    models, kind, nraw = orc.neighbours(P)
    assert len(models) <= nraw
    assert list(kind) == sorted(kind)                                          # insertions, reversals, deletions
    seen = set()
    for m, kd in zip(models, kind):
        assert (np.diag(m) == 1).all()
        key = m.tobytes()
        assert key not in seen                                                 # unique(): keep first
        seen.add(key)
        if kd == 0:
            assert (m >= P).all() and (m != P).any()                           # an insertion only adds edges
        if kd == 2:
            # a deletion removes one edge of the transitive reduction -- which is taken of the CLOSURE, so on a
            # non-closed graph the edge may not exist and the candidate is the graph itself (R/mnems_low.r:1070-1084)
            assert (m <= P).all() and int((P != m).sum()) <= 1


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 5), st.integers(3, 14), st.integers(0, 2 ** 31 - 1))
def test_greedy_search_ends_in_a_local_optimum(S, E, seed):
    """nem(search = 'greedy'): the trace is increasing, and the returned score is the score of the returned graph."""
    rng = np.random.default_rng(seed)
    R = np.round(rng.normal(size=(E, S)) * 3, 2)
    fit = orc.nem_greedy(R, None)
    sc, th, _ = orc.score_adj(R, fit["adj"], trans_close=True)
    assert abs(sc - fit["score"]) <= 1e-12 * max(1.0, abs(sc))                 # score of closure(adj) (R/mnems.r:362-369)
    assert np.array_equal(th, fit["theta"])
    assert fit["max_trace"] == fit["score"]                                    # hill-climb: the last accepted score is the best
    assert fit["score"] >= orc.score_adj(R, np.zeros((S, S)), trans_close=True)[0]


def test_inner_threads_do_not_change_a_single_bit():
    """The scale goldens (tests/golden/make_scale_golden.py) run one fit on several threads: OpenMP splits independent
    outputs only, so the whole EM result must be identical to the serial oracle, single and multi knock-down."""
    from mnem_b200 import simdata
    sim = simdata.make_config(dict(S=7, Egenes=9, C=2100, mw=[0.3, 0.3, 0.4], k=3, starts=2, uninform=3), seed=41)
    D = simdata.host_data(sim) + np.random.default_rng(1).normal(size=(sim["E"], sim["C"])) * 0.4
    init = simdata.init_phi(7, 3, 2, seed=3)
    m = simdata.make_multi(S=5, Egenes=6, C=700, seed=9)
    init_m = simdata.init_phi(5, 2, 2, seed=4)
    for acc in ("ld", "f64"):
        try:
            orc.set_inner_threads(1, acc)
            a = orc.mnem_em(D, sim["pert"], 7, init, complete=True, acc=acc)
            am = orc.mnem_em(m["D"], m["pert_enc"], 5, init_m, acc=acc)
            orc.set_inner_threads(5, acc)
            b = orc.mnem_em(D, sim["pert"], 7, init, complete=True, acc=acc)
            bm = orc.mnem_em(m["D"], m["pert_enc"], 5, init_m, acc=acc)
        finally:
            orc.set_inner_threads(1, acc)
        for x, y in ((a, b), (am, bm)):
            for key in ("adj", "theta", "probs", "mw", "ll", "n_iter", "n_scored", "min_margin", "min_accept_margin"):
                assert np.array_equal(x[key], y[key]), key
            for key in ("lls", "mwevo"):
                assert all(np.array_equal(p, q) for p, q in zip(x[key], y[key])), key


def test_null_component_of_the_oracle():
    """mnem(nullcomp = TRUE): the last component keeps an empty graph, theta 0 and log ratios 0 (R/mnems_low.r:607-608,
    R/mnems.r:2076-2080) and still receives mixture weight."""
    import oracle as orc
    from mnem_b200 import simdata
    sim = simdata.make_config(dict(S=5, Egenes=6, C=900, mw=[0.5, 0.5], k=2, starts=2, uninform=0), seed=13)
    D = simdata.host_data(sim)
    D[:, ::5] = np.random.default_rng(1).normal(size=D[:, ::5].shape)
    init = simdata.init_phi(5, 2, 2, seed=2)
    init3 = np.concatenate([init, np.ones_like(init[:, :1])], axis=1)        # the last graph is ignored
    r = orc.mnem_em(D, sim["pert"], 5, init3, nullcomp=True)
    for s in range(2):
        assert not r["adj"][s][2].any() and not r["theta"][s][2].any()
        assert np.all(r["probs"][s][2] == 0.0) and 1e-4 < r["mw"][s][2] < 0.9
    plain = orc.mnem_em(D, sim["pert"], 5, init3)
    assert plain["adj"][0][2].any() or plain["theta"][0][2].any()
