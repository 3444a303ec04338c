"""Host-side mirror of the reference's R entry points on top of the C ABI (include/mnemcu.h).

Function names, argument meaning and return fields follow the R package (cbg-ethz/mnem 1.23.0):
mnem() R/mnems.r:1691, nem() :52, scoreAdj() :416, getAffinity() :1390, getLL() R/mnems_low.r:540,
getProbs() :576, doMean() :871, transitive.closure/.reduction R/mnems.r:549/511, getIC() :1477, mnemk() :1560.
Everything numerical runs in libmnemcu.so on the GPU; this file only marshals arrays and applies the
reference's tiny host-side bookkeeping (labels, result assembly, getIC arithmetic).

Conventions: D is E x C (numpy, any order); `pert` gives the knocked-down S-gene of each column, either
integers 1..S or labels (sorted naturally and renamed 1..S like modData(), R/mnems_low.r:415-433); graphs are
S x S 0/1 arrays with adj[i, j] = 1 <=> edge i -> j; theta in 0..S with 0 = unattached.
"""
import ctypes as C
import re

import numpy as np

from . import _lib
from ._lib import EmOpts, EmStats, check, dp, i32p, load_library, u8p

MAX_S = 32


# ---- marshalling ---------------------------------------------------------------------------------------------
def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _graphs(a, S=None):
    """[n, S, S] (or [S, S]) -> uint8 buffer, each graph column-major; returns (buffer, n, S)."""
    a = (np.asarray(a) != 0).astype(np.uint8)
    if a.ndim == 2:
        a = a[None]
    n, S1, S2 = a.shape
    if S1 != S2 or (S is not None and S1 != S):
        raise ValueError("graphs must be square S x S")
    return np.ascontiguousarray(np.transpose(a, (0, 2, 1))), n, S1


def _ungraphs(buf, n, S):
    return np.transpose(np.asarray(buf).reshape(n, S, S), (0, 2, 1)).copy()


def _natural_key(s):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(s))]


def mod_data(pert):
    """modData()/getSgenes(): labels -> integers 1..S in natural order.  Returns (pert int32[C], S, labels).

    Labels with "_" ("1_3": several knock-downs in one cell) cannot be expressed as one integer per cell: use
    get_rho() and the Rho path (Context(D, Rho=...) does this automatically)."""
    p = np.asarray(pert)
    if p.dtype.kind in "iu":
        S = int(p.max())
        if p.min() < 1:
            raise ValueError("integer perturbation labels must be 1..S")
        return np.ascontiguousarray(p, dtype=np.int32), S, [str(i) for i in range(1, S + 1)]
    labs = sorted(set(str(x) for x in p), key=_natural_key)
    if any("_" in x for x in labs):
        raise ValueError("multi-knock-down labels: use get_rho() / Context(D, pert) which takes the Rho path")
    idx = {x: i + 1 for i, x in enumerate(labs)}
    return np.array([idx[str(x)] for x in p], dtype=np.int32), len(labs), labs


def get_sgenes(pert):
    """getSgenes() R/mnems_low.r:1015-1023: the distinct S-gene names in the column labels, "_"-separated
    combinations split, natural order."""
    names = set()
    for x in np.asarray(pert).ravel():
        names.update(t for t in str(x).split("_") if t != "")
    return sorted(names, key=_natural_key)


def get_rho(pert, sgenes=None):
    """getRho() R/mnems_low.r:308-324: S x C 0/1 matrix, Rho[s, c] = 1 <=> S-gene s is knocked down in cell c.
    Column labels are S-gene names joined by "_" ("3", "1_3", ...)."""
    labs = [str(x) for x in np.asarray(pert).ravel()]
    sg = list(sgenes) if sgenes is not None else get_sgenes(labs)
    idx = {s: i for i, s in enumerate(sg)}
    Rho = np.zeros((len(sg), len(labs)))
    for c, x in enumerate(labs):
        for t in x.split("_"):
            if t != "":
                Rho[idx[t], c] = 1.0
    return Rho, sg


def launch_count():
    return int(load_library().mnemcu_launch_count())


def device_count():
    n = C.c_int(0)
    check(load_library().mnemcu_device_count(C.byref(n)))
    return n.value


# ---- device-resident data ----------------------------------------------------------------------------------
class Context:
    """The expression matrix resident in HBM (mnemcu_ctx).  Create once, reuse for every call on the same data."""

    def __init__(self, D=None, pert=None, device=0, Rho=None, _handle=None, _dims=None):
        self._lib = load_library()
        self._h = _lib.c_ctx()
        if _handle is not None:
            self._h, (self.E, self.C, self.S) = _handle, _dims
            return
        D = _f(D)
        if Rho is None and pert is not None and np.asarray(pert).dtype.kind not in "iu" and \
                any("_" in str(x) for x in np.asarray(pert).ravel()):
            Rho, self.labels = get_rho(pert)                  # multi-knock-down labels: the Rho path
        if Rho is not None:
            Rho = _f(Rho)
            S = Rho.shape[0]
            if Rho.shape[1] != D.shape[1]:
                raise ValueError("Rho must be S x C with one column per column of D")
            if not hasattr(self, "labels"):
                self.labels = [str(i) for i in range(1, S + 1)]
            self.E, self.C, self.S = D.shape[0], D.shape[1], S
            check(self._lib.mnemcu_ctx_create_rho(dp(D), self.E, self.C, dp(Rho), S, int(device), C.byref(self._h)))
            return
        pert, S, self.labels = mod_data(pert)
        if D.shape[1] != len(pert):
            raise ValueError("pert must have one entry per column of D")
        self.E, self.C, self.S = D.shape[0], D.shape[1], S
        check(self._lib.mnemcu_ctx_create(dp(D), self.E, self.C, i32p(pert), S, int(device), C.byref(self._h)))

    @classmethod
    def synthetic(cls, sim, device=0):
        """Data generated on the device from a simdata.make_config() description."""
        lib = load_library()
        h = _lib.c_ctx()
        pert = np.ascontiguousarray(sim["pert"], dtype=np.int32)
        comp = np.ascontiguousarray(sim["comp"], dtype=np.uint8)
        phi, n, S = _graphs(sim["phi_true"])
        th = np.ascontiguousarray(sim["theta_true"], dtype=np.int32)
        check(lib.mnemcu_ctx_create_synthetic(sim["E"], len(pert), i32p(pert), S, u8p(comp), n, u8p(phi), i32p(th),
                                              C.c_uint64(sim["seed"]), int(device), C.byref(h)))
        return cls(_handle=h, _dims=(sim["E"], len(pert), S))

    @classmethod
    def from_device(cls, ptr, E, C_, pert, device=0):
        """D already resident on `device` at address `ptr` (column-major E x C FP64, e.g. a torch tensor's
        data_ptr() of shape [C, E]); the buffer is only read."""
        lib = load_library()
        h = _lib.c_ctx()
        pert, S, _ = mod_data(pert)
        if len(pert) != C_:
            raise ValueError("pert must have one entry per column of D")
        check(lib.mnemcu_ctx_create_device(C.c_void_p(int(ptr)), int(E), int(C_), i32p(pert), S, int(device), C.byref(h)))
        return cls(_handle=h, _dims=(int(E), int(C_), S))

    def read_data(self, c0=0, nc=None):
        nc = self.C - c0 if nc is None else nc
        out = np.zeros((self.E, nc), order="F")
        check(self._lib.mnemcu_ctx_read_data(self._h, c0, nc, dp(out)))
        return out

    def close(self):
        if self._h:
            self._lib.mnemcu_ctx_destroy(self._h)
            self._h = _lib.c_ctx()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ctx(D, pert, Rho=None):
    if isinstance(D, Context):
        return D, False
    return Context(D, pert, Rho=Rho), True


# ---- src/mm.cpp ------------------------------------------------------------------------------------------------
def transClose_W(x):
    b, n, S = _graphs(x)
    check(load_library().mnemcu_trans_close_w(u8p(b), S, n))
    r = _ungraphs(b, n, S)
    return r[0] if np.asarray(x).ndim == 2 else r


def _uv(u, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(u, dtype=np.int32), (n,)))


def transClose_Ins(x, u, v):
    b, n, S = _graphs(x)
    check(load_library().mnemcu_trans_close_ins(u8p(b), S, n, i32p(_uv(u, n)), i32p(_uv(v, n))))
    r = _ungraphs(b, n, S)
    return r[0] if np.asarray(x).ndim == 2 else r


def transClose_Del(x, u, v):
    b, n, S = _graphs(x)
    check(load_library().mnemcu_trans_close_del(u8p(b), S, n, i32p(_uv(u, n)), i32p(_uv(v, n))))
    r = _ungraphs(b, n, S)
    return r[0] if np.asarray(x).ndim == 2 else r


def maxCol_row(x):
    x = _f(x)
    out = np.zeros(x.shape[0], dtype=np.int32)
    check(load_library().mnemcu_max_col_row(dp(x), x.shape[0], x.shape[1], i32p(out)))
    return out


def eigenMapMatMult(A, B):
    A, B = _f(A), _f(B)
    if A.shape[1] != B.shape[0]:
        raise ValueError("non-conformable arguments")
    out = np.zeros((A.shape[0], B.shape[1]), order="F")
    check(load_library().mnemcu_matmul(dp(A), dp(B), A.shape[0], A.shape[1], B.shape[1], dp(out)))
    return out


# ---- graphs ------------------------------------------------------------------------------------------------------
def transitive_closure(g, u=None, v=None):
    """transitive.closure -> mytc (R/mnems_low.r:1199-1215): diagonal set to 1, then W / Ins / Del."""
    g = (np.asarray(g) != 0).astype(np.uint8).copy()
    idx = np.arange(g.shape[-1])
    g[..., idx, idx] = 1
    if u is None or v is None:
        return transClose_W(g)
    if g.ndim != 2:
        raise ValueError("u, v form takes one graph")
    return transClose_Ins(g, u, v) if g[u - 1, v - 1] == 1 else transClose_Del(g, u, v)


def transitive_reduction(g):
    b, n, S = _graphs(g)
    out = np.zeros_like(b)
    check(load_library().mnemcu_trans_reduce(u8p(b), S, n, u8p(out)))
    r = _ungraphs(out, n, S)
    return r[0] if np.asarray(g).ndim == 2 else r


def neighbours(P):
    """c(get.ins.fast(P), get.rev.tc(P), get.del.tc(P)) without unique(): (models [n,S,S], kind [n])."""
    b, _, S = _graphs(P)
    cap = max(1, 3 * S * (S - 1))
    models = np.zeros(cap * S * S, dtype=np.uint8)
    kind = np.zeros(cap, dtype=np.int8)
    n = C.c_int32(0)
    check(load_library().mnemcu_neighbours(u8p(b), S, u8p(models), kind.ctypes.data_as(_lib.c_i8p), C.byref(n)))
    return _ungraphs(models[: n.value * S * S], n.value, S), kind[: n.value].copy()


# ---- responsibilities / likelihood ------------------------------------------------------------------------------
def getAffinity(x, affinity=0, norm=True, logtype=2, mw=None, data=None, complete=False):
    if not norm:
        raise NotImplementedError("norm = FALSE returns x unchanged in the reference; not routed to the GPU")
    x = _f(x)
    k, Cn = x.shape
    out = np.zeros((k, Cn), order="F")
    m = _f(mw) if mw is not None else None
    check(load_library().mnemcu_get_affinity(dp(x), k, Cn, dp(m), int(affinity), int(bool(complete)), float(logtype), dp(out)))
    return out


def getLL(x, logtype=2, mw=None, data=None, complete=False):
    x = _f(x)
    k, Cn = x.shape
    m = _f(mw) if mw is not None else None
    ll = C.c_double(0.0)
    check(load_library().mnemcu_get_ll(dp(x), k, Cn, dp(m), int(bool(complete)), float(logtype), C.byref(ll)))
    return ll.value


# ---- M-step -----------------------------------------------------------------------------------------------------
def doMean(D, weights=None, pert=None, Rho=None):
    """doMean on the Rho path: E x S matrix (or [n, E, S] for an [n, C] weight matrix).  With multi-knock-down cells
    (Rho given, or "_" labels in pert) a cell's weighted column is split evenly over its S-genes (:897-902)."""
    cx, own = _ctx(D, pert, Rho)
    try:
        w = None
        n = 1
        if weights is not None:
            w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
            single = w.ndim == 1
            w = w.reshape(-1, cx.C)
            n = w.shape[0]
        out = np.zeros((n, cx.S, cx.E))             # per block: E x S column-major
        check(load_library().mnemcu_do_mean(cx._h, dp(w), n, dp(out)))
        R = np.transpose(out, (0, 2, 1))
        return R[0] if (weights is None or single) else R
    finally:
        if own:
            cx.close()


def scoreAdj(D, adj, pert=None, trans_close=True, dotopo=True, marginal=False, logtype=2):
    """scoreAdj(D, adj, method='llr', marginal, logtype): list(score, subtopo, subweights).

    D may be the raw E x C data (collapsed with doMean first: sum_c D[e,c] adj[pert(c), j] = (R %*% adj)[e, j]) or
    an already collapsed E x S matrix when pert is None.  adj may be one graph or a stack [n, S, S]."""
    R = _f(D) if pert is None and not isinstance(D, Context) else doMean(D, None, pert)
    R = _f(R)
    E, S = R.shape
    b, n, S2 = _graphs(adj)
    if S2 != S:
        raise ValueError("adj must be S x S with S = number of S-genes")
    score = np.zeros(n)
    theta = np.zeros((n, E), dtype=np.int32)
    W = np.zeros((n, S, E))
    if marginal:
        check(load_library().mnemcu_score_adj_marginal(dp(R), E, S, u8p(b), n, int(bool(trans_close)), float(logtype),
                                                       dp(score), i32p(theta), dp(W)))
    else:
        check(load_library().mnemcu_score_adj(dp(R), E, S, u8p(b), n, int(bool(trans_close)), dp(score), i32p(theta), dp(W)))
    W = np.transpose(W, (0, 2, 1))
    if np.asarray(adj).ndim == 2:
        return dict(score=float(score[0]), subtopo=theta[0] if dotopo else None, subweights=W[0])
    return dict(score=score, subtopo=theta if dotopo else None, subweights=W)


def nem_greedy(R, start=None, marginal=False, logtype=2):
    """Batched nem(search='greedy', marginal, logtype) on collapsed matrices R [n, E, S] (or [E, S]); start [n, S, S]
    or None."""
    R = np.asarray(R, dtype=np.float64)
    single = R.ndim == 2
    if single:
        R = R[None]
    n, E, S = R.shape
    Rb = np.ascontiguousarray(np.transpose(R, (0, 2, 1)))
    sb = None
    if start is not None:
        sb, ns, _ = _graphs(start, S)
        if ns != n:
            raise ValueError("one start graph per search")
    adj = np.zeros(n * S * S, dtype=np.uint8)
    theta = np.zeros((n, E), dtype=np.int32)
    score, mtr = np.zeros(n), np.zeros(n)
    nit = np.zeros(n, dtype=np.int32)
    nsc = np.zeros(n, dtype=np.int64)
    if marginal:
        check(load_library().mnemcu_nem_greedy_marginal(dp(Rb), E, S, u8p(sb), n, float(logtype), u8p(adj), i32p(theta),
                                                        dp(score), dp(mtr), i32p(nit), nsc.ctypes.data_as(_lib.c_i64p)))
    else:
        check(load_library().mnemcu_nem_greedy(dp(Rb), E, S, u8p(sb), n, u8p(adj), i32p(theta), dp(score), dp(mtr),
                                               i32p(nit), nsc.ctypes.data_as(_lib.c_i64p)))
    out = dict(adj=_ungraphs(adj, n, S), subtopo=theta, score=score, max_trace=mtr, n_iter=nit, n_scored=nsc)
    if single:
        out = {k_: v[0] for k_, v in out.items()}
    return out


def nem(D, pert=None, start=None, weights=None, Rho=None, marginal=False, logtype=2):
    """nem(D, search='greedy', start, weights, domean=TRUE, Rho, marginal, logtype): list(adj, score, subtopo, ...)."""
    R = doMean(D, weights, pert, Rho)
    r = nem_greedy(R, None if start is None else np.asarray(start)[None] if np.asarray(start).ndim == 2 else start,
                   marginal=marginal, logtype=logtype)
    sw = scoreAdj(R, r["adj"])["subweights"]
    return dict(adj=r["adj"], score=float(r["score"]), scores_max=float(r["max_trace"]), subtopo=r["subtopo"],
                subweights=sw, n_iter=int(r["n_iter"]), n_scored=int(r["n_scored"]))


# ---- E-step -----------------------------------------------------------------------------------------------------
def getProbs(probs, k, data, res, pert=None, mw=None, complete=False, logtype=2, Rho=None):
    """getProbs(probs, k, data, res, ..., Rho): list(probs, subtopoX, mw, ll).

    res: list of k dicts with 'adj' and optionally 'subtopo' (all or none), like the reference's res list."""
    cx, own = _ctx(data, pert, Rho)
    try:
        phi = np.stack([np.asarray(r["adj"]) for r in res])
        has = [r.get("subtopo") is not None for r in res]
        if any(has) and not all(has):
            raise ValueError("either every component carries subtopo or none does")
        th = np.ascontiguousarray(np.stack([r["subtopo"] for r in res]), dtype=np.int32) if all(has) else None
        pb, kk, S = _graphs(phi, cx.S)
        if kk != k:
            raise ValueError("res must hold k components")
        L_in = _f(probs)
        if L_in.shape != (k, cx.C):
            raise ValueError("probs must be k x C")
        L_out = np.zeros((k, cx.C), order="F")
        mw_out = np.zeros(k)
        ll = np.zeros(1)
        th_used = np.zeros((k, cx.E), dtype=np.int32)
        m = _f(mw) if mw is not None else None
        check(load_library().mnemcu_get_probs(cx._h, 1, k, u8p(pb), i32p(th), dp(L_in), dp(m), int(bool(complete)),
                                              float(logtype), dp(L_out), dp(mw_out), dp(ll), i32p(th_used)))
        return dict(probs=L_out, subtopoX=None, mw=mw_out, ll=float(ll[0]), theta_used=th_used)
    finally:
        if own:
            cx.close()


# ---- EM ---------------------------------------------------------------------------------------------------------
def sample_init(S, k, starts, seed=0):
    """Initial graphs with the distribution of initComps(linets=FALSE) -> sampleRndNetwork(DAG=TRUE)
    (R/mnems_low.r:326-347, 60-94).  R's RNG stream cannot be reproduced outside R; pass phi= for seeded parity."""
    from .simdata import init_phi
    return init_phi(S, k, starts, seed)


def parse_tape(words, k, E, S):
    """Decision tape of mnemcu_em_run (record layout: include/mnemcu.h, mnemcu_em_opts::tape) -> {start: dict}.

    Per start: init_theta [T][k][E], init_best [T], searches {(iter, comp, restart): [n][S] uint32 row masks},
    pick [n_iter][k], theta [n_iter][k][E], accept / improved [n_iter], best_at (sorted counts), n_iter."""
    w = np.asarray(words, dtype=np.int32)
    out = {}
    kE = k * E

    def st(s):
        return out.setdefault(int(s), dict(init_theta={}, init_best={}, searches={}, pick={}, theta={}, accept={},
                                           improved={}, best_at=[], n_iter=None))
    i = 0
    while i < len(w):
        code, s = int(w[i]), int(w[i + 1])
        d = st(s)
        if code == 1:
            d["init_theta"][int(w[i + 2])] = w[i + 3:i + 3 + kE].reshape(k, E).copy()
            i += 3 + kE
        elif code == 6:
            d["init_best"][int(w[i + 2])] = int(w[i + 3])
            i += 4
        elif code == 2:
            it, comp, rst, n = (int(x) for x in w[i + 2:i + 6])
            d["searches"][(it, comp, rst)] = w[i + 6:i + 6 + n * S].view(np.uint32).reshape(n, S).copy()
            i += 6 + n * S
        elif code == 3:
            d["pick"][(int(w[i + 2]), int(w[i + 3]))] = int(w[i + 4])
            i += 5
        elif code == 4:
            d["theta"][int(w[i + 2])] = w[i + 3:i + 3 + kE].reshape(k, E).copy()
            i += 3 + kE
        elif code == 5:
            d["accept"][int(w[i + 2])] = int(w[i + 3])
            d["improved"][int(w[i + 2])] = int(w[i + 4])
            i += 5
        elif code == 7:
            d["best_at"].append(int(w[i + 2]))
            i += 3
        elif code == 8:
            d["n_iter"] = int(w[i + 2])
            i += 3
        else:
            raise ValueError("decision tape: unknown record code %d at word %d" % (code, i))
    return out


def em_run(cx, init_phi, complete=False, logtype=2.0, batch=0, max_trace=256, max_iter_guard=0, want_probs=True,
           next_start=None, affinity=0, tape=False, probs="all", mw=None, nullcomp=False, theta=None):
    """mnemcu_em_run on a Context: all starts, returns per-start arrays + winner + stats.

    next_start: optional callable returning the next start index to fit (or -1 when none is left) -- the ticket
    function of a queue shared by several ranks (mnem_b200.dist.StartQueue).  Starts fitted elsewhere come back with
    n_iter = -1 and ll = -inf."""
    init_phi = np.asarray(init_phi)
    starts, k = init_phi.shape[:2]
    S, E, Cn = cx.S, cx.E, cx.C
    mw0 = None
    if mw is not None:                     # mnem(mw =): initial mixture weights of every start (R/mnems.r:1771-1773)
        mw0 = np.ascontiguousarray(mw, dtype=np.float64)
        if mw0.shape != (k,):
            raise ValueError("mw must hold k weights")
    th_fixed = None
    if theta is not None:                  # mnem(theta =): fixed attachments, one [k, E] block per start (R/mnems.r:1890-1892)
        th_fixed = np.ascontiguousarray(theta, dtype=np.int32)
        if th_fixed.shape != (starts, k, E):
            raise ValueError("theta must be [starts, k, E]")
    pb, _, _ = _graphs(init_phi.reshape(starts * k, S, S), S)
    adj = np.zeros(starts * k * S * S, dtype=np.uint8)
    theta = np.zeros((starts, k, E), dtype=np.int32)
    if probs not in ("all", "winner"):
        raise ValueError("probs must be 'all' or 'winner'")
    winner_only = probs == "winner"       # ONE k x C block: that of the best start this call fitted (what mnem() returns)
    probs = np.zeros((1 if winner_only else starts, Cn, k)) if want_probs else None
    mw = np.zeros((starts, k))
    ll = np.zeros(starts)
    nit = np.zeros(starts, dtype=np.int32)
    tr = [np.zeros((starts, max_trace)) for _ in range(4)]
    best = C.c_int32(-1)
    st = EmStats()
    cb_err = []

    def _ticket(_user):
        try:
            return int(next_start())
        except BaseException as ex:            # never let an exception unwind through the C frames
            cb_err.append(ex)
            return -1

    cb = _lib.NEXT_START_FN(_ticket) if next_start is not None else _lib.NEXT_START_FN()
    tape_buf = None
    if tape:        # audit runs: capacity for 64 EM iterations of 2k searches with S*S/2 moves each, per start (checked by the library)
        tape_buf = np.zeros(starts * (2 * k * 64 * (S * S // 2 + 8) * S + (64 + 100) * (k * E + 16)) + 1024, dtype=np.int32)
    o = EmOpts(int(bool(complete)), float(logtype), int(batch), int(max_trace), int(max_iter_guard), cb, None, int(affinity),
               tape_buf.ctypes.data_as(C.POINTER(C.c_int32)) if tape else None, tape_buf.size if tape else 0,
               1 if winner_only else 0, dp(mw0) if mw0 is not None else None, int(bool(nullcomp)),
               i32p(th_fixed) if th_fixed is not None else None)
    check(load_library().mnemcu_em_run(cx._h, starts, k, u8p(pb), C.byref(o), u8p(adj), i32p(theta), dp(probs), dp(mw),
                                       dp(ll), i32p(nit), dp(tr[0]), dp(tr[1]), dp(tr[2]), dp(tr[3]), C.byref(best),
                                       C.byref(st)))
    if cb_err:
        raise cb_err[0]
    stats = {f: getattr(st, f) for f, _ in EmStats._fields_}
    extra = dict(tape=parse_tape(tape_buf[:st.tape_words], k, E, S), tape_words=tape_buf[:st.tape_words].copy()) if tape else {}
    return dict(**extra, best=best.value, adj=_ungraphs(adj, starts * k, S).reshape(starts, k, S, S), theta=theta,
                probs=np.transpose(probs, (0, 2, 1)) if want_probs else None, mw=mw, ll=ll, n_iter=nit,
                lls=[tr[0][s, :max(0, min(nit[s], max_trace))].copy() for s in range(starts)],
                phievo=[tr[1][s, :max(0, min(nit[s], max_trace))].copy() for s in range(starts)],
                thetaevo=[tr[2][s, :max(0, min(nit[s], max_trace))].copy() for s in range(starts)],
                mwevo=[tr[3][s, :max(0, min(nit[s], max_trace))].copy() for s in range(starts)], stats=stats)


def mnem(D, pert=None, k=None, starts=3, phi=None, complete=False, logtype=2, seed=0, batch=0, max_iter_guard=0,
         Rho=None, affinity=0, mw=None, nullcomp=False, theta=None):
    """mnem(D, inference='em', search='greedy', k, starts, phi, complete, logtype, Rho, affinity, mw, nullcomp, theta).

    theta: [starts, k, E] fixed E-gene attachments (0..S), mnem(theta =): the graphs are searched, theta never moves.

    nullcomp = True adds a null component as component k + 1 (R/mnems.r:1832-1834): empty graph, theta 0, log ratios 0.

    Returns the fields of the reference's `mnem` object that the hot path produces: limits (per start), comp
    (phi reduced, theta), mw (lambda0, R/mnems.r:2298-2303), probs (k x C log ratios of the winner), lls, phievo,
    thetaevo, mwevo, ll, complete, plus best_start and the device-side stats."""
    cx, own = _ctx(D, pert, Rho)
    try:
        if phi is not None:
            phi = np.asarray(phi)
            starts, k = phi.shape[:2]                                 # R/mnems.r:1765-1767
        elif k is None:
            raise NotImplementedError("learnk (R/mnems_low.r:437) is out of scope: pass k or phi")
        else:
            phi = sample_init(cx.S, k, starts, seed)
        if nullcomp:                                                  # k <- k + 1, res1[[k]]$adj <- adj * 0  (:1832-1834, 1894-1898)
            phi = np.concatenate([phi, np.zeros_like(phi[:, :1])], axis=1)
            k = phi.shape[1]
        if k == 1:
            return _mnem_k1(cx, phi[0, 0], complete, logtype)
        r = em_run(cx, phi, complete=complete, logtype=logtype, batch=batch, max_iter_guard=max_iter_guard, affinity=affinity,
                   mw=mw, nullcomp=nullcomp, theta=theta)
        b = r["best"]
        limits = [dict(adj=r["adj"][s], theta=r["theta"][s], probs=r["probs"][s], mw=r["mw"][s], ll=r["ll"][s],
                       lls=r["lls"][s], phievo=r["phievo"][s], thetaevo=r["thetaevo"][s], mwevo=r["mwevo"][s])
                  for s in range(starts)]
        probs = r["probs"][b]
        post = getAffinity(probs, mw=r["mw"][b], logtype=logtype, complete=complete, affinity=affinity)   # R/mnems.r:2298-2303
        lam = post.sum(1)
        comp = [dict(phi=r["adj"][b][i], theta=r["theta"][b][i]) for i in range(k)]
        return dict(limits=limits, comp=comp, mw=lam / lam.sum(), probs=probs, lls=r["lls"][b], phievo=r["phievo"][b],
                    thetaevo=r["thetaevo"][b], mwevo=r["mwevo"][b], ll=float(r["ll"][b]), complete=complete,
                    best_start=b, stats=r["stats"], n_iter=r["n_iter"])
    finally:
        if own:
            cx.close()


def _mnem_k1(cx, start, complete, logtype):
    """k == 1 branch of mnem(), R/mnems.r:1835-1876: one nem() on all cells, then getProbs with mw = 1."""
    fit = nem(cx, start=start)
    res = [dict(adj=fit["adj"], subtopo=fit["subtopo"])]
    pr = getProbs(np.zeros((1, cx.C)), 1, cx, res, mw=np.ones(1), complete=complete, logtype=logtype)
    lim = dict(adj=fit["adj"][None], theta=fit["subtopo"][None], probs=pr["probs"], mw=np.ones(1), ll=pr["ll"],
               lls=np.array([pr["ll"]]))
    return dict(limits=[lim], comp=[dict(phi=fit["adj"], theta=fit["subtopo"])], mw=np.ones(1), probs=pr["probs"],
                lls=lim["lls"], ll=pr["ll"], complete=complete, best_start=0, stats={}, n_iter=np.array([0]))


def getIC(x, n_cells, man=False, degree=4, logtype=2, pen=2):
    """getIC(x, degree=4) R/mnems.r:1477-1529 (useF = FALSE): host arithmetic on the fitted object."""
    comp = x["comp"]
    k = len(comp)
    if degree != 5:
        fpar = 0
        for c in comp:
            t = transitive_closure(c["phi"]) if degree > 2 else transitive_reduction(c["phi"])
            t = np.array(t, copy=True)
            np.fill_diagonal(t, 0)
            fpar += int((t != 0).sum())
        if degree > 1 and degree != 3:
            fpar += k * len(comp[0]["theta"])
    else:
        fpar = (comp[0]["phi"].size + len(comp[0]["theta"])) * k
    if degree > 0:
        fpar += k - 1
    LL = np.max(x["ll"]) * np.log(logtype)
    return (pen if man else np.log(n_cells)) * fpar - 2 * LL


def bootstrap(x, D, pert=None, size=1000, p=1.0, logtype=2, complete=False, seed=0, chunk=256, part=(0, 1), fit_fn=None):
    """bootstrap(x, size, p) R/mnems.r:2348-2369: edge support of every component graph.

    For component i the data are weighted by its responsibilities and collapsed (doMean inside nem()), the E-gene
    rows are resampled with replacement `size` times, nem() is run from the empty graph on every resample -- all of
    them as ONE batch of device searches -- and the transitive closures are averaged.  The row samples come from a
    seeded NumPy generator (R's sample() stream cannot be reproduced outside R); returns (support [k,S,S], idx).

    part = (r, n): fit only the resamples r, r + n, r + 2n, ... and return their closure COUNTS divided by `size` -- the
    share of rank r of n; the shares add up to the full support (mnem_b200.dist.bootstrap_sharded all-reduces them).
    fit_fn(R_batch) -> closed graphs [n, S, S] replaces the device searches in CPU tests."""
    cx, own = _ctx(D, pert)
    try:
        k = len(x["comp"])
        post = getAffinity(x["probs"], mw=x["mw"], logtype=logtype, complete=complete)
        R = doMean(cx, post)                                   # [k, E, S]: rowsums of t(t(data) * post[i, ]) per S-gene
        R = R[None] if R.ndim == 2 else R
        return _bootstrap_core(R, k, cx.S, size, p, seed, chunk, part, fit_fn)
    finally:
        if own:
            cx.close()


def _bootstrap_core(R, k, S, size, p, seed, chunk, part, fit_fn):
    E = R.shape[1]
    ne = int(np.ceil(p * E))
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, E, size=(k, size, ne))               # sample(seq_len(nrow), ceiling(p*nrow), replace = TRUE)
    mine = np.arange(part[0], size, part[1])
    support = np.zeros((k, S, S))
    if fit_fn is None:
        def fit_fn(Rb):
            return transitive_closure(nem_greedy(Rb, None)["adj"])      # [n, ne, S] searches from the empty start
    for i in range(k):
        for j0 in range(0, len(mine), chunk):
            sel = idx[i, mine[j0:j0 + chunk]]
            if len(sel):
                support[i] += np.asarray(fit_fn(R[i][sel])).sum(0)
    return support / size, idx


def mnemk(D, pert=None, ks=(1, 2, 3, 4, 5), **kw):
    """mnemk(): fit every k in ks, keep the minimum information criterion (R/mnems.r:1560-1574)."""
    cx, own = _ctx(D, pert)
    try:
        fits, ics, lls = {}, {}, {}
        for k in ks:
            fits[k] = mnem(cx, k=k, **kw)
            ics[k] = getIC(fits[k], cx.C)
            lls[k] = fits[k]["ll"]
        kb = min(ics, key=lambda q: ics[q])
        return dict(best=fits[kb], ics=ics, lls=lls, k=kb)
    finally:
        if own:
            cx.close()
