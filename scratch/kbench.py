"""Kernel timing sweep: E-step and M-step statistics kernels alone, CUDA events, over pairs per launch."""
import sys, json, ctypes as C
import numpy as np
sys.path.insert(0, '.')
import mnem_b200 as mb
from mnem_b200 import simdata, _lib
from mnem_b200.api import _graphs
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
batches = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 3, 4]
nrep = int(sys.argv[3]) if len(sys.argv) > 3 else 6
sim = simdata.make_config(cfg)
k, S, E, Cn = sim["k"], sim["S"], sim["E"], sim["C"]
lib = mb.load_library()
rng = np.random.default_rng(0)
with mb.Context.synthetic(sim) as cx:
    for B in batches:
        phi = simdata.init_phi(S, k, B, seed=3).reshape(B * k, S, S)
        pb, _, _ = _graphs(phi)
        theta = rng.integers(0, S + 1, (B * k, E)).astype(np.int32)
        ms = np.zeros(nrep); by = C.c_double(0)
        _lib.check(lib.mnemcu_bench_estep(cx._h, B, k, _lib.u8p(pb), _lib.i32p(theta), nrep, _lib.dp(ms), C.byref(by)))
        e = ms[2:].mean()
        ms2 = np.zeros(nrep); by2 = C.c_double(0)
        _lib.check(lib.mnemcu_bench_mstats(cx._h, B, k, nrep, _lib.dp(ms2), C.byref(by2)))
        m = ms2[2:].mean()
        print(json.dumps(dict(cfg=cfg, batch=B, pairs=B * k, estep_ms=e, estep_gbs=by.value / e / 1e6, mstats_ms=m,
                              mstats_gbs=by2.value / m / 1e6, estep_all=[round(x, 3) for x in ms], mstats_all=[round(x, 3) for x in ms2])))
