"""Pins the CPU oracle against everything the reference itself ships as a known answer.

* data/app.rda: 9 fitted objects carry (probs, mw, ll); ll must equal getLL(probs, mw)
  (R/mnems_low.r:540-552) -- fixture tests/golden/app_rda_golden.npz, made by make_app_rda_golden.py.
* transitive.closure roxygen example R/mnems.r:546-548.
* hard-clustering property of inst/unitTests/test_hardClustering.R:1-13.
"""
import os

import numpy as np
import pytest

import oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "app_rda_golden.npz"))
KEYS = sorted(k[:-6] for k in G.files if k.endswith("_probs"))


@pytest.mark.parametrize("key", KEYS)
def test_getll_known_answers(key):
    probs, mw, ll = G[key + "_probs"], G[key + "_mw"], G[key + "_ll"]
    got = orc.get_ll(probs, mw, complete=False, logtype=2.0)
    assert abs(got - ll.max()) <= 2e-16 * abs(got), (got, ll)
    # the double-accumulator build stays within the north-star tolerance of the long-double one
    got64 = orc.get_ll(probs, mw, acc="f64")
    assert abs(got64 - got) <= 1e-12 * abs(got)


@pytest.mark.parametrize("key", KEYS)
def test_affinity_columns_sum_to_one(key):
    probs, mw = G[key + "_probs"], G[key + "_mw"]
    g = orc.get_affinity(probs, mw)
    assert g.shape == probs.shape
    np.testing.assert_allclose(g.sum(0), 1.0, rtol=0, atol=4e-16 * probs.shape[0])
    if probs.shape[0] == 1:
        assert (g == 1).all()


@pytest.mark.parametrize("key", KEYS)
def test_hard_clustering_property(key):
    """test_hardClustering.R: affinity=1 equals the one-hot argmax of the soft responsibilities."""
    probs, mw = G[key + "_probs"], G[key + "_mw"]
    if probs.shape[0] == 1:
        pytest.skip("k = 1")
    soft = orc.get_affinity(probs, mw)
    hard = orc.get_affinity(probs, mw, affinity=1)
    want = (soft == soft.max(0, keepdims=True)).astype(float)
    # cells whose top-2 responsibilities are within rounding can legitimately differ: none in the fixtures
    assert (hard == want).all()


def test_closure_roxygen_example():
    g = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], dtype=float).reshape(3, 3, order="F")   # R/mnems.r:547
    tc = orc.mytc(g)
    assert (tc == np.triu(np.ones((3, 3)))).all()
    tr = orc.transitive_reduction(g)
    assert (tr == g).all()


def test_mm_cpp_hand_cases():
    # transClose_Ins: x[i,j] |= x[i,u] && x[v,j] (src/mm.cpp:50-65).  chain 1->2, 3->4, insert 2->3
    x = np.eye(4)
    x[0, 1] = x[2, 3] = 1
    x[1, 2] = 1
    y = orc.trans_close_ins(x, 2, 3)
    want = np.triu(np.ones((4, 4)))
    assert (y == want).all()
    # transClose_Del only re-derives entry (u,v) (src/mm.cpp:32-48)
    x = np.triu(np.ones((3, 3)))
    x[0, 2] = 0
    assert orc.trans_close_del(x, 1, 3)[0, 2] == 1
    x[0, 1] = 0
    assert orc.trans_close_del(x, 1, 3)[0, 2] == 0
    # maxCol_row: first strict maximum, 1-based (src/mm.cpp:66-84)
    m = np.array([[1.0, 3.0, 3.0], [-1.0, -2.0, -1.0], [0.0, 0.0, 0.0]])
    assert orc.max_col_row(m).tolist() == [2, 1, 1]
    a = np.arange(6.0).reshape(2, 3)
    b = np.arange(12.0).reshape(3, 4)
    assert (orc.matmul(a, b) == a @ b).all()


def test_phi_fixtures_closure_reduction_roundtrip():
    """closure(reduction(phi)) == closure(phi) for every shipped component graph (all are DAG-like)."""
    n = 0
    for k in G.files:
        if "_phi" in k:
            phi = G[k]
            clo = orc.mytc(phi)
            red = orc.transitive_reduction(phi)
            assert (np.diag(red) == 0).all()
            assert (orc.mytc(red) == clo).all()
            assert red.sum() <= (clo - np.eye(len(clo))).sum()
            n += 1
    assert n == 45
