"""ctypes front-end of the CPU parity oracle (oracle/mnem_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (mnem_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

c_dp = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_i32p = C.POINTER(C.c_int32)


class NemInfo(C.Structure):
    _fields_ = [("score", C.c_double), ("max_trace", C.c_double), ("n_iter", C.c_int), ("n_scored", C.c_long),
                ("min_margin", C.c_double), ("n_raw", C.c_long), ("n_dec", C.c_long), ("n_dec_tight", C.c_long)]


class EmOpts(C.Structure):
    _fields_ = [("max_trace", C.c_int), ("complete", C.c_int), ("logtype", C.c_double), ("faithful_cost", C.c_int),
                ("max_iter_guard", C.c_int), ("noise_mode", C.c_int), ("affinity", C.c_int), ("mw0", c_dp), ("nullcomp", C.c_int), ("theta0", c_i32p)]


class EmInfo(C.Structure):
    _fields_ = [("ll", C.c_double), ("best_ll", C.c_double), ("n_iter", C.c_int), ("n_scored", C.c_long),
                ("n_greedy_iter", C.c_long), ("min_margin", C.c_double), ("n_raw", C.c_long),
                ("min_accept_margin", C.c_double), ("n_dec", C.c_long), ("n_dec_tight", C.c_long)]


def build():
    """Compile the oracle with the committed Makefile (gcc only)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib(acc="ld"):
    """acc='ld': long-double accumulators (R on x86-64); acc='f64': double accumulators."""
    if acc not in _LIBS:
        name = "libmnem_oracle.so" if acc == "ld" else "libmnem_oracle_f64.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        _LIBS[acc] = C.CDLL(path)
        _LIBS[acc].orc_get_ll.restype = C.c_double
        _LIBS[acc].orc_score_adj.restype = C.c_double
        _LIBS[acc].orc_score_adj_m.restype = C.c_double
        _LIBS[acc].orc_get_probs.restype = C.c_double
    return _LIBS[acc]


def set_inner_threads(n, acc="ld"):
    """Threads INSIDE one fit (cells / E-gene rows / candidates; bit-identical results for any n).  For fits at
    BASELINE scale; the default 1 keeps every loop serial.  Do not combine with mnem_em(nthreads > 1)."""
    lib(acc).orc_set_inner_threads(int(n))


def _d(a):
    return a.ctypes.data_as(c_dp)


def _u8(a):
    return a.ctypes.data_as(c_u8p)


def _i32(a):
    return a.ctypes.data_as(c_i32p)


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _g(a):
    """graph(s) -> uint8 Fortran-ordered copy, last two dims S x S."""
    return np.asfortranarray((np.asarray(a) != 0).astype(np.uint8))


def _graphs(a):
    """stack of graphs [n, S, S] -> contiguous buffer where each graph is column-major."""
    a = (np.asarray(a) != 0).astype(np.uint8)
    if a.ndim == 2:
        a = a[None]
    return np.ascontiguousarray(np.transpose(a, (0, 2, 1)))  # [n][col][row] == col-major per graph


def _ungraphs(buf, n, S):
    return np.transpose(buf.reshape(n, S, S), (0, 2, 1)).copy()


# ---- src/mm.cpp ------------------------------------------------------------------------------------------
def trans_close_w(x, acc="ld"):
    S = x.shape[0]
    b = _graphs(x)
    lib(acc).orc_trans_close_w(_u8(b), S)
    return _ungraphs(b, 1, S)[0]


def trans_close_ins(x, u, v, acc="ld"):
    S = x.shape[0]
    b = _graphs(x)
    lib(acc).orc_trans_close_ins(_u8(b), S, int(u), int(v))
    return _ungraphs(b, 1, S)[0]


def trans_close_del(x, u, v, acc="ld"):
    S = x.shape[0]
    b = _graphs(x)
    lib(acc).orc_trans_close_del(_u8(b), S, int(u), int(v))
    return _ungraphs(b, 1, S)[0]


def max_col_row(x, acc="ld"):
    x = _f(x)
    out = np.zeros(x.shape[0], dtype=np.int32)
    lib(acc).orc_max_col_row(_d(x), x.shape[0], x.shape[1], _i32(out))
    return out


def matmul(A, B, acc="ld"):
    A, B = _f(A), _f(B)
    out = np.zeros((A.shape[0], B.shape[1]), order="F")
    lib(acc).orc_matmul(_d(A), _d(B), A.shape[0], A.shape[1], B.shape[1], _d(out))
    return out


# ---- graphs ----------------------------------------------------------------------------------------------
def mytc(x, u=None, v=None, acc="ld"):
    S = x.shape[0]
    b = _graphs(x)
    if u is None or v is None:
        lib(acc).orc_mytc(_u8(b), S)
    else:
        lib(acc).orc_mytc_uv(_u8(b), S, int(u), int(v))
    return _ungraphs(b, 1, S)[0]


def transitive_reduction(g, acc="ld"):
    S = g.shape[0]
    b = _graphs(g)
    out = np.zeros_like(b)
    lib(acc).orc_trans_reduce(_u8(b), S, _u8(out))
    return _ungraphs(out, 1, S)[0]


def neighbours(P, acc="ld"):
    """unique(c(get.ins.fast, get.rev.tc, get.del.tc)(P)) -> (models [n,S,S], kind [n], n_raw)."""
    S = P.shape[0]
    b = _graphs(P)
    models = np.zeros(3 * S * S * S * S + S * S, dtype=np.uint8)
    kind = np.zeros(3 * S * S + 1, dtype=np.int8)
    nraw = C.c_int(0)
    n = lib(acc).orc_neighbours(_u8(b), S, _u8(models), kind.ctypes.data_as(C.POINTER(C.c_int8)), C.byref(nraw))
    return _ungraphs(models[: n * S * S], n, S), kind[:n].copy(), nraw.value


# ---- likelihood ------------------------------------------------------------------------------------------
def get_affinity(L, mw=None, affinity=0, complete=False, logtype=2.0, acc="ld"):
    L = _f(L)
    k, Cn = L.shape
    out = np.zeros((k, Cn), order="F")
    mwp = _d(_f(mw)) if mw is not None else None
    lib(acc).orc_get_affinity(_d(L), k, Cn, mwp, int(affinity), int(bool(complete)), C.c_double(logtype), _d(out))
    return out


def get_ll(L, mw=None, complete=False, logtype=2.0, acc="ld"):
    L = _f(L)
    k, Cn = L.shape
    mwp = _d(_f(mw)) if mw is not None else None
    return lib(acc).orc_get_ll(_d(L), k, Cn, mwp, int(bool(complete)), C.c_double(logtype))


def mw_update(L, mw, complete=False, logtype=2.0, na_guard=False, acc="ld"):
    L = _f(L)
    k, Cn = L.shape
    m = _f(mw).copy()
    lib(acc).orc_mw_update(_d(L), k, Cn, _d(m), int(bool(complete)), C.c_double(logtype), int(na_guard))
    return m


# ---- M-step ----------------------------------------------------------------------------------------------
def do_mean(D, pert, S, weights=None, acc="ld"):
    D = _f(D)
    E, Cn = D.shape
    pert = np.ascontiguousarray(pert, dtype=np.int32)
    R = np.zeros((E, S), order="F")
    wp = _d(_f(weights)) if weights is not None else None
    lib(acc).orc_do_mean(_d(D), E, Cn, _i32(pert), S, wp, _d(R))
    return R


def score_adj(R, adj, trans_close=True, marginal=False, logtype=2.0, acc="ld"):
    """scoreAdj(D=R, adj, marginal, dotopo=TRUE) -> (score, theta 0..S, W)."""
    R = _f(R)
    E, S = R.shape
    b = _graphs(adj)
    theta = np.zeros(E, dtype=np.int32)
    W = np.zeros((E, S), order="F")
    sc = lib(acc).orc_score_adj_m(_d(R), E, S, _u8(b), int(bool(trans_close)), int(bool(marginal)), C.c_double(logtype),
                                  _i32(theta), _d(W))
    return sc, theta, W


def nem_greedy(R, start=None, marginal=False, logtype=2.0, acc="ld"):
    """nem(search='greedy', marginal) on collapsed data -> dict(adj reduced, theta, score, max_trace, n_iter, n_scored)."""
    R = _f(R)
    E, S = R.shape
    sb = _graphs(start) if start is not None else None
    adj = np.zeros(S * S, dtype=np.uint8)
    theta = np.zeros(E, dtype=np.int32)
    info = NemInfo()
    lib(acc).orc_nem_greedy_m(_d(R), E, S, _u8(sb) if sb is not None else None, int(bool(marginal)), C.c_double(logtype),
                              _u8(adj), _i32(theta), C.byref(info))
    return dict(adj=_ungraphs(adj, 1, S)[0], theta=theta, score=info.score, max_trace=info.max_trace,
                n_iter=info.n_iter, n_scored=info.n_scored, min_margin=info.min_margin, n_raw=info.n_raw,
                n_dec=info.n_dec, n_dec_tight=info.n_dec_tight)


# ---- E-step ----------------------------------------------------------------------------------------------
def get_probs(D, pert, S, phi, theta, L_in, mw, complete=False, logtype=2.0, faithful_cost=False, acc="ld"):
    """getProbs -> dict(probs k x C, mw, ll, theta_used k x E)."""
    D = _f(D)
    E, Cn = D.shape
    pert = np.ascontiguousarray(pert, dtype=np.int32)
    pb = _graphs(phi)
    k = pb.shape[0]
    th = np.ascontiguousarray(theta, dtype=np.int32) if theta is not None else None
    L_in = _f(L_in)
    L_out = np.zeros((k, Cn), order="F")
    mw_out = np.zeros(k)
    th_used = np.zeros((k, E), dtype=np.int32)
    mwp = _d(_f(mw)) if mw is not None else None
    ll = lib(acc).orc_get_probs(_d(D), E, Cn, _i32(pert), S, k, _u8(pb), _i32(th) if th is not None else None,
                                _d(L_in), mwp, int(bool(complete)), C.c_double(logtype), int(faithful_cost),
                                _d(L_out), _d(mw_out), _i32(th_used))
    return dict(probs=L_out, mw=mw_out, ll=ll, theta_used=th_used)


# ---- EM --------------------------------------------------------------------------------------------------
def mnem_em(D, pert, S, init_phi, complete=False, logtype=2.0, nthreads=1, faithful_cost=False, max_trace=256,
            max_iter_guard=0, noise_mode=0, affinity=0, acc="ld", mw=None, nullcomp=False, theta=None):
    """mnem(D, phi=init_phi) EM phase.  init_phi: [starts, k, S, S].  Returns per-start limits and the winner.
    theta: optional [starts, k, E] fixed attachments (mnem(theta =)): the starts are then fitted one call each."""
    if theta is not None:
        theta = np.ascontiguousarray(theta, dtype=np.int32)
        parts = []
        for s_ in range(len(init_phi)):
            parts.append(_mnem_em_one_theta(D, pert, S, np.asarray(init_phi)[s_:s_ + 1], theta[s_], complete, logtype, max_trace,
                                            max_iter_guard, noise_mode, affinity, acc, mw, nullcomp))
        out = {q: (np.concatenate([p_[q] for p_ in parts]) if isinstance(parts[0][q], np.ndarray) else sum((p_[q] for p_ in parts), []))
               for q in parts[0] if q != "best"}
        best = 0
        for s_ in range(1, len(init_phi)):
            if out["ll"][best] <= out["ll"][s_]:
                best = s_
        out["best"] = best
        return out
    mw0 = np.ascontiguousarray(mw, dtype=np.float64) if mw is not None else None       # mnem(mw =), R/mnems.r:1771-1773
    D = _f(D)
    E, Cn = D.shape
    pert = np.ascontiguousarray(pert, dtype=np.int32)
    init_phi = np.asarray(init_phi)
    starts, k = init_phi.shape[:2]
    pb = _graphs(init_phi.reshape(starts * k, S, S))
    adj = np.zeros(starts * k * S * S, dtype=np.uint8)
    theta = np.zeros((starts, k, E), dtype=np.int32)
    L = np.zeros((starts, Cn, k))          # per start: k x C column-major == [C][k] C-order
    mw = np.zeros((starts, k))
    tr = [np.zeros((starts, max_trace)) for _ in range(4)]
    infos = (EmInfo * starts)()
    o = EmOpts(max_trace, int(bool(complete)), logtype, int(faithful_cost), max_iter_guard, int(noise_mode), int(affinity),
               _d(mw0) if mw0 is not None else None, int(bool(nullcomp)))
    best = lib(acc).orc_mnem_em(_d(D), E, Cn, _i32(pert), S, k, starts, _u8(pb), C.byref(o), int(nthreads), _u8(adj),
                                _i32(theta), _d(L), _d(mw), _d(tr[0]), _d(tr[1]), _d(tr[2]), _d(tr[3]), infos)
    adj = _ungraphs(adj, starts * k, S).reshape(starts, k, S, S)
    n_iter = np.array([i.n_iter for i in infos])
    res = dict(best=best, adj=adj, theta=theta, probs=np.transpose(L, (0, 2, 1)).copy(), mw=mw,
               ll=np.array([i.ll for i in infos]), n_iter=n_iter,
               n_scored=np.array([i.n_scored for i in infos]), n_greedy_iter=np.array([i.n_greedy_iter for i in infos]),
               min_margin=np.array([i.min_margin for i in infos]), n_raw=np.array([i.n_raw for i in infos]),
               min_accept_margin=np.array([i.min_accept_margin for i in infos]),
               n_dec=np.array([i.n_dec for i in infos]), n_dec_tight=np.array([i.n_dec_tight for i in infos]),
               lls=[tr[0][s, :n_iter[s]].copy() for s in range(starts)],
               phievo=[tr[1][s, :n_iter[s]].copy() for s in range(starts)],
               thetaevo=[tr[2][s, :n_iter[s]].copy() for s in range(starts)],
               mwevo=[tr[3][s, :n_iter[s]].copy() for s in range(starts)])
    return res


def simgen(sim, c0=0, nc=None, nthreads=8):
    """Cells [c0, c0+nc) of the synthetic matrix of a mnem_b200.simdata config dict (E x nc, Fortran order) from the
    oracle's own C twin of the generator (oracle/simgen.c) -- the CPU arm never loads the product library."""
    nc = sim["C"] - c0 if nc is None else nc
    out = np.zeros((sim["E"], nc), order="F")
    pert = np.ascontiguousarray(sim["pert"], dtype=np.int32)
    comp = np.ascontiguousarray(sim["comp"], dtype=np.uint8)
    phi = _graphs(sim["phi_true"])
    th = np.ascontiguousarray(sim["theta_true"], dtype=np.int32)
    lib().orc_simgen_fill(_d(out), int(sim["E"]), C.c_int64(c0), C.c_int64(nc), _i32(pert), int(sim["S"]), _u8(comp), _u8(phi),
                          _i32(th), C.c_uint64(sim["seed"]), int(nthreads))
    return out


ADV_KINDS = ("greedy_move", "greedy_stop", "restart_pick", "theta", "init_theta", "init_best", "accept", "best_or_length")


class Advice(C.Structure):
    _fields_ = [("tol", C.c_double), ("tol_ll", C.c_double), ("T", C.c_int), ("init_theta", c_i32p), ("init_best", c_i32p),
                ("n_em", C.c_int), ("move_off", C.POINTER(C.c_int64)), ("move_rows", C.POINTER(C.c_uint32)),
                ("pick", c_i32p), ("theta", c_i32p), ("accept", c_i32p), ("best_at", c_i32p),
                ("n_checked", C.c_long * 8), ("n_forced", C.c_long * 8), ("n_hard", C.c_long * 8),
                ("max_forced", C.c_double * 8), ("min_hard", C.c_double * 8), ("diverged", C.c_int),
                ("first_hard", C.c_char * 256)]


def record_start(D, pert, S, init_phi, complete=False, logtype=2.0, acc="ld", start=0, cap=1 << 24, mw=None):
    """One EM start fitted by the oracle while it writes a decision tape of its own (device record format).  Returns the
    int32 words (parse with mnem_b200.api.parse_tape) and the fit's n_iter / ll."""
    mw0 = np.ascontiguousarray(mw, dtype=np.float64) if mw is not None else None
    D = _f(D)
    E, Cn = D.shape
    pert = np.ascontiguousarray(pert, dtype=np.int32)
    init_phi = np.asarray(init_phi)
    k = init_phi.shape[0]
    buf = np.zeros(cap, dtype=np.int32)
    L_ = lib(acc)
    L_.orc_recorder_words.restype = C.c_int64
    L_.orc_set_recorder(_i32(buf), C.c_int64(cap), int(start))
    try:
        pb = _graphs(init_phi.reshape(k, S, S))
        adj = np.zeros(k * S * S, dtype=np.uint8)
        th = np.zeros((k, E), dtype=np.int32)
        L = np.zeros((Cn, k))
        mw = np.zeros(k)
        tr = [np.zeros(256) for _ in range(4)]
        info = EmInfo()
        o = EmOpts(256, int(bool(complete)), logtype, 0, 0, 0, 0, _d(mw0) if mw0 is not None else None)
        L_.orc_em_start(_d(D), E, Cn, _i32(pert), S, k, _u8(pb), C.byref(o), _u8(adj), _i32(th), _d(L), _d(mw), _d(tr[0]),
                        _d(tr[1]), _d(tr[2]), _d(tr[3]), C.byref(info))
        n = int(L_.orc_recorder_words())
    finally:
        L_.orc_set_recorder(None, C.c_int64(0), 0)
    if n > cap:
        raise RuntimeError("tape capacity %d < %d words" % (cap, n))
    return dict(words=buf[:n].copy(), n_iter=info.n_iter, ll=info.ll, adj=_ungraphs(adj, k, S), theta=th, mw=mw,
                lls=tr[0][:info.n_iter].copy())


def replay_start(D, pert, S, init_phi, tape, complete=False, logtype=2.0, tol=1e-13, tol_ll=1e-12, max_trace=256,
                 affinity=0, acc="ld", mw=None, max_iter_guard=0, nullcomp=False, theta=None):
    """One EM start replayed against the decision tape of a device fit (mnem_b200.api.parse_tape(...)[start]).

    The oracle runs the start in its own arithmetic and takes the tape's choice wherever its own decision differs by
    less than `tol` (scores, weights; relative) / `tol_ll` (log-likelihoods) -- see orc_advice in mnem_oracle.c.
    Returns the fit (like mnem_em, one start) plus report = {kind: dict(checked, forced, hard, max_forced, min_hard)},
    diverged and first_hard."""
    mw0 = np.ascontiguousarray(mw, dtype=np.float64) if mw is not None else None
    th0 = np.ascontiguousarray(theta, dtype=np.int32) if theta is not None else None      # mnem(theta =): fixed attachments
    D = _f(D)
    E, Cn = D.shape
    pert = np.ascontiguousarray(pert, dtype=np.int32)
    init_phi = np.asarray(init_phi)
    k = init_phi.shape[0]
    assert C.sizeof(Advice) == lib(acc).orc_advice_size()
    T = len(tape["init_theta"])
    n_em = int(tape["n_iter"])
    init_theta = np.ascontiguousarray(np.stack([tape["init_theta"][t] for t in range(T)] + [np.zeros((k, E), dtype=np.int32)]),
                                      dtype=np.int32)                  # (T = 0 with fixed theta: nothing to replay there)
    init_best = np.array([tape["init_best"][t] for t in range(T)] + [0], dtype=np.int32)
    off = np.zeros(n_em * k * 2 + 1, dtype=np.int64)
    rows = []
    for it in range(n_em):
        for i in range(k):
            for j in range(2):
                m = tape["searches"].get((it, i, j), np.zeros((0, S), dtype=np.uint32))      # no search: the null component
                rows.append(m.reshape(-1, S))
                q = (it * k + i) * 2 + j
                off[q + 1] = off[q] + len(m)
    rows = np.ascontiguousarray(np.concatenate(rows + [np.zeros((1, S), dtype=np.uint32)]), dtype=np.uint32)
    pick = np.array([tape["pick"].get((it, i), 0) for it in range(n_em) for i in range(k)] + [0], dtype=np.int32)
    theta = np.ascontiguousarray(np.stack([tape["theta"][it] for it in range(n_em)] + [np.zeros((k, E), dtype=np.int32)]),
                                 dtype=np.int32)
    accept = np.array([tape["accept"][it] for it in range(n_em)] + [0], dtype=np.int32)
    best_at = np.zeros(n_em + 2, dtype=np.int32)
    for c in tape["best_at"]:
        best_at[c] = 1
    adv = Advice()
    adv.tol, adv.tol_ll, adv.T, adv.n_em = tol, tol_ll, T, n_em
    adv.init_theta, adv.init_best = _i32(init_theta), _i32(init_best)
    adv.move_off, adv.move_rows = off.ctypes.data_as(C.POINTER(C.c_int64)), rows.ctypes.data_as(C.POINTER(C.c_uint32))
    adv.pick, adv.theta, adv.accept, adv.best_at = _i32(pick), _i32(theta), _i32(accept), _i32(best_at)
    pb = _graphs(init_phi.reshape(k, S, S))
    adj = np.zeros(k * S * S, dtype=np.uint8)
    th = np.zeros((k, E), dtype=np.int32)
    L = np.zeros((Cn, k))
    mw = np.zeros(k)
    tr = [np.zeros(max_trace) for _ in range(4)]
    info = EmInfo()
    o = EmOpts(max_trace, int(bool(complete)), logtype, 0, int(max_iter_guard), 0, int(affinity),
               _d(mw0) if mw0 is not None else None, int(bool(nullcomp)), _i32(th0) if th0 is not None else None)
    lib(acc).orc_em_start_advised(_d(D), E, Cn, _i32(pert), S, k, _u8(pb), C.byref(o), C.byref(adv), _u8(adj), _i32(th),
                                  _d(L), _d(mw), _d(tr[0]), _d(tr[1]), _d(tr[2]), _d(tr[3]), C.byref(info))
    n = info.n_iter
    report = {name: dict(checked=int(adv.n_checked[q]), forced=int(adv.n_forced[q]), hard=int(adv.n_hard[q]),
                         max_forced=float(adv.max_forced[q]), min_hard=float(adv.min_hard[q]))
              for q, name in enumerate(ADV_KINDS)}
    return dict(adj=_ungraphs(adj, k, S), theta=th, probs=L.T.copy(), mw=mw, ll=info.ll, n_iter=n, lls=tr[0][:n].copy(),
                phievo=tr[1][:n].copy(), thetaevo=tr[2][:n].copy(), mwevo=tr[3][:n].copy(), n_raw=info.n_raw,
                n_dec=info.n_dec, n_dec_tight=info.n_dec_tight, min_margin=info.min_margin, report=report,
                diverged=bool(adv.diverged), first_hard=adv.first_hard.decode("utf-8", "replace"))


def _mnem_em_one_theta(D, pert, S, init1, theta1, complete, logtype, max_trace, max_iter_guard, noise_mode, affinity, acc, mw,
                       nullcomp):
    """mnem_em for ONE start with fixed theta (orc_em_opts.theta0 belongs to a start)."""
    D = _f(D)
    E, Cn = D.shape
    pert = np.ascontiguousarray(pert, dtype=np.int32)
    k = init1.shape[1]
    mw0 = np.ascontiguousarray(mw, dtype=np.float64) if mw is not None else None
    th0 = np.ascontiguousarray(theta1, dtype=np.int32)
    pb = _graphs(init1.reshape(k, S, S))
    adj = np.zeros(k * S * S, dtype=np.uint8)
    theta = np.zeros((1, k, E), dtype=np.int32)
    L = np.zeros((1, Cn, k))
    mwo = np.zeros((1, k))
    tr = [np.zeros((1, max_trace)) for _ in range(4)]
    infos = (EmInfo * 1)()
    o = EmOpts(max_trace, int(bool(complete)), logtype, 0, max_iter_guard, int(noise_mode), int(affinity),
               _d(mw0) if mw0 is not None else None, int(bool(nullcomp)), _i32(th0))
    lib(acc).orc_mnem_em(_d(D), E, Cn, _i32(pert), S, k, 1, _u8(pb), C.byref(o), 1, _u8(adj), _i32(theta), _d(L), _d(mwo),
                         _d(tr[0]), _d(tr[1]), _d(tr[2]), _d(tr[3]), infos)
    n = infos[0].n_iter
    return dict(best=0, adj=_ungraphs(adj, k, S).reshape(1, k, S, S), theta=theta, probs=np.transpose(L, (0, 2, 1)).copy(), mw=mwo,
                ll=np.array([infos[0].ll]), n_iter=np.array([n]), min_margin=np.array([infos[0].min_margin]),
                min_accept_margin=np.array([infos[0].min_accept_margin]), n_raw=np.array([infos[0].n_raw]),
                lls=[tr[0][0, :n].copy()], phievo=[tr[1][0, :n].copy()], thetaevo=[tr[2][0, :n].copy()], mwevo=[tr[3][0, :n].copy()])


def final_mw(L, mw_best, complete=False, logtype=2.0, affinity=0, acc="ld"):
    L = _f(L)
    k, Cn = L.shape
    out = np.zeros(k)
    lib(acc).orc_final_mw_a(_d(L), k, Cn, _d(_f(mw_best)), int(affinity), int(bool(complete)), C.c_double(logtype), _d(out))
    return out


# ---- the reference's own src/mm.cpp, compiled unmodified (oracle/_ref/libmm_ref.so) --------------------------
_REF = {}


def ref_lib():
    """The reference's src/mm.cpp compiled verbatim against the stub headers of oracle/rstub (oracle/Makefile).  Built
    where /root/reference exists; elsewhere (the GPU box) the prebuilt oracle/_ref/libmm_ref.so is loaded.  Returns
    None when neither is available."""
    if "lib" not in _REF:
        path = os.path.join(_HERE, "_ref", "libmm_ref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/src/mm.cpp"):
            build()
        _REF["lib"] = C.CDLL(path) if os.path.exists(path) else None
    return _REF["lib"]


def _ref_graph(x):
    return np.asfortranarray(np.asarray(x, dtype=np.float64)).copy(order="F")


def ref_trans_close_w(x):
    """transClose_W(x) of src/mm.cpp:15-30 on an R-style double matrix (returned as uint8 0/1)."""
    b = _ref_graph(x)
    ref_lib().mmref_trans_close_w(_d(b), b.shape[0])
    return b.astype(np.uint8)


def ref_trans_close_ins(x, u, v):
    b = _ref_graph(x)
    ref_lib().mmref_trans_close_ins(_d(b), b.shape[0], int(u), int(v))
    return b.astype(np.uint8)


def ref_trans_close_del(x, u, v):
    b = _ref_graph(x)
    ref_lib().mmref_trans_close_del(_d(b), b.shape[0], int(u), int(v))
    return b.astype(np.uint8)


def ref_max_col_row(x):
    x = _f(x).copy(order="F")
    out = np.zeros(x.shape[0], dtype=np.int32)
    ref_lib().mmref_max_col_row(_d(x), x.shape[0], x.shape[1], _i32(out))
    return out


def ref_matmul(A, B):
    A, B = _f(A).copy(order="F"), _f(B).copy(order="F")
    out = np.zeros((A.shape[0], B.shape[1]), order="F")
    ref_lib().mmref_matmul(_d(A), _d(B), A.shape[0], A.shape[1], B.shape[1], _d(out))
    return out
