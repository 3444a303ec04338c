"""The synthetic input generator: NumPy reference vs the library's threaded host generator (bit-identical)."""
import numpy as np

from mnem_b200 import simdata


def test_host_generators_agree_bit_for_bit():
    sim = simdata.make_config(dict(S=6, Egenes=5, C=500, mw=[0.3, 0.7], k=2, starts=2, uninform=3), seed=21)
    D = simdata.host_data(sim)
    Dn = simdata.host_data_native(sim, nthreads=3)
    assert D.shape == (33, sim["C"]) and np.array_equal(D, Dn)
    part = simdata.host_data_native(sim, c0=100, nc=57, nthreads=2)
    assert np.array_equal(part, D[:, 100:157])


def test_generator_statistics():
    sim = simdata.make_config(dict(S=8, Egenes=6, C=4000, mw=[0.5, 0.5], k=2, starts=1, uninform=0), seed=2)
    D = simdata.host_data(sim)
    noise = D - np.sign(np.round(D * 65536) / 65536 - 0) * 0  # noqa: F841  (values are multiples of 2^-16)
    # full-precision values: hardly any is a multiple of 2^-16 (sums over them are order-dependent, as with real data)
    assert np.mean(D * 65536 == np.round(D * 65536)) < 1e-3
    comp, pert, th, phi = sim["comp"].astype(int), sim["pert"] - 1, sim["theta_true"], sim["phi_true"]
    b = phi[comp[None, :], pert[None, :], th[comp].T - 1]
    res = D - np.where(b != 0, 1.0, -1.0)
    assert abs(res.mean()) < 0.01 and abs(res.std() - 1.0) < 0.01
    assert abs(np.mean(sim["comp"] == 0) - 0.5) < 0.01
    cnt = np.bincount(sim["pert"], minlength=9)[1:]
    assert cnt.min() > 0.8 * cnt.mean()


def test_configs_have_the_survey_shapes():
    want = {1: (5, 10, 1000), 2: (10, 100, 10002), 3: (20, 1000, 100000), 4: (24, 1000, 200016), 5: (30, 2000, 1000020)}
    for c, (S, E, Cn) in want.items():
        cfg = simdata.CONFIGS[c]
        assert cfg["S"] == S and cfg["S"] * cfg["Egenes"] + cfg["uninform"] == E
    sim = simdata.make_config(2)
    assert abs(sim["C"] - 10002) <= 3 and sim["E"] == 100


def test_init_phi_is_closed_lower_triangular():
    ph = simdata.init_phi(7, 3, 5, seed=4)
    assert ph.shape == (5, 3, 7, 7)
    for g in ph.reshape(-1, 7, 7):
        assert (np.triu(g) == 0).all()
        c = simdata._closure(g.copy())
        np.fill_diagonal(c, 0)
        assert (c == g).all()
