"""A/B of the FP32-filter + FP64-rescore search against the all-FP64 search: results must be bit-identical."""
import os, subprocess, sys, json
code = r'''
import sys, json, time, numpy as np
sys.path.insert(0, ".")
import mnem_b200 as mb
rng = np.random.default_rng(0)
S, E, n = 30, 2000, 40
def planted():
    a = np.triu((rng.random((S, S)) < 0.3).astype(np.uint8), 1)
    t = mb.transitive_closure(a)
    p = rng.permutation(S); t = t[np.ix_(p, p)]
    th = rng.integers(1, S + 1, E)
    return np.where(t[:, th - 1].T > 0, 1.0, -1.0) * 4 + rng.normal(size=(E, S)) * 3
R = np.stack([planted() for _ in range(n)])
R[3, :, 5] = R[3, :, 4]            # indistinguishable S-genes: exact ties
R[4] = np.round(R[4])              # coarse values: many exact ties
out = mb.nem_greedy(R, None)
t0 = time.time(); out = mb.nem_greedy(R, None); dt = time.time() - t0
import hashlib
h = hashlib.sha256()
for k in ("adj", "subtopo", "score", "max_trace", "n_iter", "n_scored"):
    h.update(np.ascontiguousarray(out[k]).tobytes())
print(json.dumps(dict(seconds=dt, digest=h.hexdigest(), iters=int(out["n_iter"].sum()))))
'''
res = {}
for mode in ("0", "1"):
    env = dict(os.environ, MNEMCU_EXACT_SCORES=mode)
    o = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(mode, o.stdout.strip(), o.stderr.strip()[-500:])
    res[mode] = json.loads(o.stdout.strip().splitlines()[-1])
print("IDENTICAL" if res["0"]["digest"] == res["1"]["digest"] else "MISMATCH", res["1"]["seconds"] / res["0"]["seconds"])
