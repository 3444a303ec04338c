"""simData-style synthetic inputs (restated from R/mnems.r:2989-3160, README.md:36-40) for tests and bench.py.

The reference draws everything from R's RNG, which cannot be reproduced outside R.  Here the small structural
tables (component graphs, E-gene attachments, cell labels) come from a seeded NumPy PCG64 generator and the
E x C noise comes from a counter-based integer hash, so that the device generator
(mnemcu_ctx_create_synthetic), the host generator (mnemcu_synthetic_fill_host) and the NumPy reference below
produce bit-identical matrices for any sub-range of cells.

  D[e, c] = (bin[e, c] - 0.5) / 0.5 + noise(seed, e + E*c)                                   README.md:38-39
  bin[e, c] = closure(Nem_i)[pert(c), theta_i(e)]  for the component i of cell c             R/mnems.r:3084-3138
  noise = (sum of twelve 16-bit uniforms - 6*65535) / 65536   (Irwin-Hall(12): mean 0, variance 1.0000)
        + (53-bit uniform - 0.5) / 65536                      (full-precision mantissas: sums are order-dependent)
"""
import numpy as np

# BASELINE.json configs: Egenes is per S-gene, uninform rows are appended
CONFIGS = {
    1: dict(S=5, Egenes=2, C=1000, mw=[0.4, 0.6], k=2, starts=10, uninform=0, complete=False),
    2: dict(S=10, Egenes=10, C=10000, mw=[1 / 3] * 3, k=3, starts=20, uninform=0, complete=False),
    3: dict(S=20, Egenes=50, C=100000, mw=[0.2] * 5, k=5, starts=50, uninform=0, complete=True),
    4: dict(S=24, Egenes=41, C=200000, mw=[1 / 3] * 3, k=3, starts=3, uninform=16, complete=True),
    5: dict(S=30, Egenes=66, C=1000000, mw=[0.2] * 5, k=5, starts=100, uninform=20, complete=True),
}


def _closure(a):
    a = (a != 0).astype(np.uint8)
    np.fill_diagonal(a, 1)
    n = len(a)
    for k in range(n):
        a |= np.outer(a[:, k], a[k, :])
    return a


def make_config(cfg, seed=None):
    """cfg: int key of CONFIGS or a dict with S, Egenes, C, mw, k, starts, uninform."""
    if isinstance(cfg, int):
        seed = 1000 + cfg if seed is None else seed
        cfg = CONFIGS[cfg]
    seed = 0 if seed is None else seed
    rng = np.random.default_rng(seed)
    S, Eg, nC, mw, un = cfg["S"], cfg["Egenes"], cfg["C"], list(cfg["mw"]), cfg.get("uninform", 0)
    n_comp = len(mw)
    E = S * Eg + un
    phi_true = np.zeros((n_comp, S, S), dtype=np.uint8)
    theta_true = np.zeros((n_comp, E), dtype=np.int32)
    perts, comps = [], []
    reps = -(-nC // S)
    for i in range(n_comp):
        a = np.triu((rng.random((S, S)) < 0.3).astype(np.uint8), 1)         # strictly upper => DAG
        perm = rng.permutation(S)                                            # random relabelling, R/mnems.r:3040-3043
        a = a[np.ix_(perm, perm)]
        phi_true[i] = _closure(a)
        pool = np.tile(np.arange(1, S + 1, dtype=np.int32), reps)           # rep(seq_len(S), reps2)  :3086
        n_i = int(np.ceil(mw[i] * len(pool) - 1e-9))                         # ceiling(mw[i]*ncol)     :3090-3092
        perts.append(pool[rng.permutation(len(pool))[:n_i]])
        comps.append(np.full(n_i, i, dtype=np.uint8))
        th = np.repeat(np.arange(1, S + 1, dtype=np.int32), Eg)             # rep(seq_len(S), each=Egenes) :3133
        theta_true[i, : S * Eg] = th[rng.permutation(S * Eg)]               # eorder :3136-3138 (uninform rows: 0)
    pert = np.concatenate(perts).astype(np.int32)
    comp = np.concatenate(comps)
    out = dict(cfg)
    out.update(S=S, E=E, C=len(pert), pert=pert, comp=comp, phi_true=phi_true, theta_true=theta_true,
               seed=int(seed) * 7919 + 17, n_comp=n_comp)
    return out


_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def noise(seed, idx):
    """NumPy twin of syn_noise() in mnem_b200/csrc/ctx.cu."""
    idx = idx.astype(np.uint64)
    tot = np.zeros(idx.shape, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for j in range(3):
            w = _mix64(np.uint64(seed) + _G * (np.uint64(3) * idx + np.uint64(j + 1)))
            tot += (w & np.uint64(0xFFFF)) + ((w >> np.uint64(16)) & np.uint64(0xFFFF)) + \
                   ((w >> np.uint64(32)) & np.uint64(0xFFFF)) + (w >> np.uint64(48))
        w4 = _mix64((np.uint64(seed) ^ np.uint64(0xA0761D6478BD642F)) + np.uint64(0xE7037ED1A0B428DB) * (idx + np.uint64(1)))
    coarse = (tot.astype(np.float64) - 393210.0) * (1.0 / 65536.0)
    fine = ((w4 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5) * (1.0 / 65536.0)
    return coarse + fine


def host_data(sim, c0=0, nc=None):
    """NumPy reference generator: E x nc block of D starting at cell c0 (Fortran order)."""
    E, S = sim["E"], sim["S"]
    nc = sim["C"] - c0 if nc is None else nc
    cells = np.arange(c0, c0 + nc, dtype=np.uint64)
    e = np.arange(E, dtype=np.uint64)
    idx = e[:, None] + np.uint64(E) * cells[None, :]
    comp = sim["comp"][c0:c0 + nc].astype(np.int64)
    pert = sim["pert"][c0:c0 + nc].astype(np.int64) - 1
    th = sim["theta_true"][comp].T                                           # E x nc
    phi = sim["phi_true"]
    b = phi[comp[None, :], pert[None, :], np.maximum(th - 1, 0)] != 0
    with np.errstate(over="ignore"):
        coin = (_mix64(np.uint64(sim["seed"]) ^ np.uint64(0xD1B54A32D192ED03) ^
                       (idx * np.uint64(0x2545F4914F6CDD1D))) & np.uint64(1)) != 0
    b = np.where(th > 0, b, coin)
    D = np.where(b, 1.0, -1.0) + noise(sim["seed"], idx)
    return np.asfortranarray(D)


def host_data_native(sim, c0=0, nc=None, nthreads=8):
    """Same matrix through the library's threaded host generator (for large bench inputs)."""
    import ctypes as C

    from . import _lib
    from .api import _graphs
    nc = sim["C"] - c0 if nc is None else nc
    out = np.zeros((sim["E"], nc), order="F")
    pert = np.ascontiguousarray(sim["pert"], dtype=np.int32)
    comp = np.ascontiguousarray(sim["comp"], dtype=np.uint8)
    phi, n, S = _graphs(sim["phi_true"])
    th = np.ascontiguousarray(sim["theta_true"], dtype=np.int32)
    _lib.check(_lib.load_library().mnemcu_synthetic_fill_host(_lib.dp(out), sim["E"], c0, nc, _lib.i32p(pert), S,
                                                              _lib.u8p(comp), n, _lib.u8p(phi), _lib.i32p(th),
                                                              C.c_uint64(sim["seed"]), nthreads))
    return out


def init_phi(S, k, starts, seed=0):
    """[starts, k, S, S] initial graphs with the distribution of sampleRndNetwork(seq_len(S), DAG=TRUE)
    (R/mnems_low.r:60-94): scale-free out-degree (gamma 2.5), edges only to lower-numbered S-genes, closed, diag 0."""
    rng = np.random.default_rng(seed)
    deg = np.arange(S + 1, dtype=float)
    degprob = np.ones(S + 1)
    degprob[1:] = deg[1:] ** -2.5
    degprob /= degprob.sum()
    out = np.zeros((starts, k, S, S), dtype=np.uint8)
    for s in range(starts):
        for i in range(k):
            a = np.eye(S, dtype=np.uint8)
            for r in range(S):
                od = rng.choice(S + 1, p=degprob)
                if od > 0 and r > 0:
                    idx0 = np.arange(r)                                       # S[i,] == 0 & seq_len(n) < i
                    pick = rng.choice(idx0, size=min(od, len(idx0)), replace=True)
                    a[r, pick] = 1
            a = _closure(a)
            np.fill_diagonal(a, 0)
            out[s, i] = a
    return out


def make_multi(S=5, Egenes=4, C=600, mw=(0.4, 0.6), multi=(0.25, 0.1), uninform=0, seed=0):
    """Small simData(multi = c(p2, p3))-style problem with double / triple knock-downs (R/mnems.r:3099-3132):
    a fraction p2 of the cells carries two knocked-down S-genes, p3 three; a cell shows the effect of an E-gene when
    ANY of its S-genes is an ancestor-or-self of the E-gene's attachment (the OR of R/mnems_low.r:613-616).
    Host-side NumPy only (tests): returns dict(D, labels, Rho, masks, pert_enc, S, E, C, k, phi_true, theta_true)
    where pert_enc is the oracle's encoding (s for a single knock-down, -mask otherwise)."""
    rng = np.random.default_rng(seed)
    k = len(mw)
    E = S * Egenes + uninform
    phi = np.zeros((k, S, S), dtype=np.uint8)
    theta = np.zeros((k, E), dtype=np.int32)
    for i in range(k):
        a = np.triu((rng.random((S, S)) < 0.3).astype(np.uint8), 1)
        perm = rng.permutation(S)
        phi[i] = _closure(a[np.ix_(perm, perm)])
        th = np.repeat(np.arange(1, S + 1, dtype=np.int32), Egenes)
        theta[i, : S * Egenes] = th[rng.permutation(S * Egenes)]
    comp = rng.choice(k, size=C, p=np.asarray(mw) / np.sum(mw))
    n_ko = rng.choice([1, 2, 3], size=C, p=[1 - multi[0] - multi[1], multi[0], multi[1]])
    masks = np.zeros(C, dtype=np.uint32)
    labels = []
    for c in range(C):
        genes = np.sort(rng.choice(S, size=min(n_ko[c], S), replace=False))
        for g in genes:
            masks[c] |= np.uint32(1) << np.uint32(g)
        labels.append("_".join(str(g + 1) for g in genes))
    Rho = np.zeros((S, C))
    for s in range(S):
        Rho[s] = (masks >> np.uint32(s)) & 1
    D = np.zeros((E, C))
    for c in range(C):
        eff = (Rho[:, c] @ phi[comp[c]]) > 0                                  # [S] which attachments are hit
        att = theta[comp[c]]
        hit = np.where(att > 0, eff[np.maximum(att, 1) - 1], rng.random(E) < 0.5)
        D[:, c] = np.where(hit, 1.0, -1.0)
    D += rng.normal(size=(E, C))
    single = np.array([bin(int(m)).count("1") == 1 for m in masks])
    pert_enc = np.where(single, np.log2(np.maximum(masks, 1)).astype(np.int64) + 1, -masks.astype(np.int64)).astype(np.int32)
    return dict(D=np.asfortranarray(D), labels=labels, Rho=Rho, masks=masks, pert_enc=pert_enc, S=S, E=E, C=C, k=k,
                phi_true=phi, theta_true=theta, comp=comp)
