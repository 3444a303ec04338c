#!/usr/bin/env python
"""Records the decision tape of a device fit at BASELINE scale (run on a B200) so that the certificate can be checked
WITHOUT a GPU: tests/test_replay_committed_tape.py regenerates the same synthetic matrix on the CPU and replays the tape
through the oracle.  Writes gpurun_out/gpu_tape_<case>_s<start>.npz (copy it to tests/golden/).

  python scripts/save_tape.py cfg3_full 0"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_scale_golden as msg  # noqa: E402
import mnem_b200 as mb  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_full"
    start = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    sim, init = msg.build_case(name)
    with mb.Context.synthetic(sim) as cx:
        r = mb.em_run(cx, init[start:start + 1], complete=sim["complete"], tape=True)
    idx = msg.sample_index(sim["C"])
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "gpu_tape_%s_s%d.npz" % (name, start)), case=name, start=start,
                        tape_words=r["tape_words"], adj=r["adj"][0], theta=r["theta"][0], ll=r["ll"][0], mw=r["mw"][0],
                        n_iter=r["n_iter"][0], lls=r["lls"][0], phievo=r["phievo"][0], thetaevo=r["thetaevo"][0],
                        probs_idx=idx, probs_sample=r["probs"][0][:, idx], init=init[start],
                        version=mb.load_library().mnemcu_version())
    print("saved", name, start, "n_iter", r["n_iter"][0], "tape words", len(r["tape_words"]))


if __name__ == "__main__":
    main()
