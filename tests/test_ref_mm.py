"""src/mm.cpp of the reference, compiled UNMODIFIED (oracle/_ref/libmm_ref.so, recipe oracle/Makefile, stub R/Rcpp/Eigen
headers oracle/rstub/), against the oracle's restatement (CPU) and against the CUDA library (gpu).

This pins rows a9 (transClose_W / _Ins / _Del, src/mm.cpp:15-65) and a11 (maxCol_row, :66-84) to the reference's own
compiled code.  eigenMapMatMult (:8-13) compiles too, but its product comes from the stub's restatement of Eigen (a
third-party dependency that is not in /root/reference), so for a7 it only checks the marshalling."""
import os

import numpy as np
import pytest

import oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "app_rda_golden.npz"))
needs_ref = pytest.mark.skipif(orc.ref_lib() is None, reason="oracle/_ref/libmm_ref.so not built (no /root/reference here)")


def graphs(seed, n=40):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        S = int(rng.integers(1, 33))
        a = (rng.random((S, S)) < rng.uniform(0.02, 0.5)).astype(np.uint8)
        if i % 2 == 0:
            a = np.triu(a, 1)
            p = rng.permutation(S)
            a = a[np.ix_(p, p)]
        if i % 3:
            np.fill_diagonal(a, 1)
        out.append(a)
    return out + [G[k] for k in G.files if "_phi" in k][:12]


@needs_ref
def test_oracle_closure_routines_equal_the_compiled_reference():
    rng = np.random.default_rng(0)
    for g in graphs(1):
        S = len(g)
        assert np.array_equal(orc.trans_close_w(g), orc.ref_trans_close_w(g))
        for _ in range(4):
            u, v = int(rng.integers(1, S + 1)), int(rng.integers(1, S + 1))
            assert np.array_equal(orc.trans_close_ins(g, u, v), orc.ref_trans_close_ins(g, u, v)), (S, u, v)
            assert np.array_equal(orc.trans_close_del(g, u, v), orc.ref_trans_close_del(g, u, v)), (S, u, v)


@needs_ref
def test_oracle_max_col_row_and_matmul_equal_the_compiled_reference():
    rng = np.random.default_rng(2)
    for _ in range(30):
        nr, nc = int(rng.integers(1, 70)), int(rng.integers(1, 34))
        x = rng.normal(size=(nr, nc)).round(1)                    # rounded: ties between columns occur
        x[rng.random((nr, nc)) < 0.05] = -np.inf
        x[rng.random((nr, nc)) < 0.03] = np.nan
        assert np.array_equal(orc.max_col_row(x), orc.ref_max_col_row(x))
        a, b = rng.normal(size=(nr, nc)), rng.normal(size=(nc, int(rng.integers(1, 20))))
        assert np.array_equal(orc.matmul(a, b), orc.ref_matmul(a, b))


def test_fitted_graphs_of_app_rda_are_fixed_points_of_the_reduction():
    """The 45 component graphs shipped in data/app.rda are outputs of transitive.reduction (R/mnems.r:374), so
    reduction(phi) == phi is a reference-held known answer for row a10 (R/mnems.r:511-533)."""
    n = 0
    for k in G.files:
        if "_phi" in k:
            assert np.array_equal(orc.transitive_reduction(G[k]), (G[k] != 0).astype(np.uint8)), k
            n += 1
    assert n == 45


def test_getic_arithmetic_on_the_shipped_fits():
    """getIC (R/mnems.r:1477-1529, degree = 4) replayed on the shipped (phi, theta, ll) of every k: the host-side getIC of
    the package (GPU closure inside) must equal this plain NumPy evaluation.  The shipped objects do not record which
    k their authors selected, so this pins the arithmetic, not a selection."""
    for ds in G["__datasets__"]:
        for k in range(1, 6):
            pre = "%s_k%d_" % (ds, k)
            if pre + "ll" not in G.files:
                continue
            phis = [G[pre + "phi%d" % i] for i in range(k)]
            E = len(G[pre + "theta0"])
            fpar = sum(int((orc.mytc(p) != 0).sum()) - len(p) for p in phis) + k * E + (k - 1)
            ic = np.log(1000) * fpar - 2 * G[pre + "ll"].max() * np.log(2.0)
            assert np.isfinite(ic)


# ---- the CUDA library against the compiled reference ------------------------------------------------------
@pytest.mark.gpu
@needs_ref
def test_cuda_closure_routines_equal_the_compiled_reference():
    import mnem_b200 as mb
    rng = np.random.default_rng(3)
    for g in graphs(4):
        S = len(g)
        assert np.array_equal(mb.transClose_W(g), orc.ref_trans_close_w(g))
        for _ in range(4):
            u, v = int(rng.integers(1, S + 1)), int(rng.integers(1, S + 1))
            assert np.array_equal(mb.transClose_Ins(g, u, v), orc.ref_trans_close_ins(g, u, v)), (S, u, v)
            assert np.array_equal(mb.transClose_Del(g, u, v), orc.ref_trans_close_del(g, u, v)), (S, u, v)


@pytest.mark.gpu
@needs_ref
def test_cuda_max_col_row_and_matmul_equal_the_compiled_reference():
    import mnem_b200 as mb
    rng = np.random.default_rng(5)
    for _ in range(20):
        nr, nc = int(rng.integers(1, 3000)), int(rng.integers(1, 34))
        x = rng.normal(size=(nr, nc)).round(1)
        assert np.array_equal(mb.maxCol_row(x), orc.ref_max_col_row(x))
        a, b = rng.normal(size=(nr, nc)), rng.normal(size=(nc, int(rng.integers(1, 33))))
        assert np.array_equal(mb.eigenMapMatMult(a, b), orc.ref_matmul(a, b))


@pytest.mark.gpu
def test_cuda_reduction_fixed_points_of_app_rda():
    import mnem_b200 as mb
    phis = [G[k] for k in G.files if "_phi" in k]
    for p in phis:
        assert np.array_equal(mb.transitive_reduction(p), (p != 0).astype(np.uint8))


@pytest.mark.gpu
def test_cuda_getic_on_the_shipped_fits():
    import mnem_b200 as mb
    for ds in G["__datasets__"]:
        ics = {}
        for k in range(1, 6):
            pre = "%s_k%d_" % (ds, k)
            if pre + "ll" not in G.files:
                continue
            comp = [dict(phi=G[pre + "phi%d" % i], theta=G[pre + "theta%d" % i]) for i in range(k)]
            E = len(comp[0]["theta"])
            got = mb.getIC(dict(comp=comp, ll=G[pre + "ll"]), 1000)
            fpar = sum(int((orc.mytc(c["phi"]) != 0).sum()) - len(c["phi"]) for c in comp) + k * E + (k - 1)
            want = np.log(1000) * fpar - 2 * G[pre + "ll"].max() * np.log(2.0)
            assert abs(got - want) <= 1e-12 * abs(want)
            ics[k] = got
        assert len(ics) >= 1
