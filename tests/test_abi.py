"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/mnemcu.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mnem_b200 as mb
from mnem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mnemcu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mnemcu_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    decl = _declared()
    assert len(decl) >= 25
    assert sorted(_lib.SIGNATURES) == decl, set(decl) ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert mb.load_library().mnemcu_version() >= 100


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under mnem_b200/ may reference it."""
    pkg = os.path.join(ROOT, "mnem_b200")
    for dp_, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp_, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "mnem_oracle" not in txt, f


@pytest.mark.skipif(mb.api.load_library() is None, reason="library missing")
def test_compute_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(mb.MnemcuError):
        mb.getLL(np.zeros((2, 8)))
    with pytest.raises(mb.MnemcuError):
        mb.Context(np.zeros((4, 8)), np.array([1, 2] * 4))
    with pytest.raises(mb.MnemcuError):
        mb.transitive_closure(np.eye(3))


def test_argument_validation_is_reported_not_crashed():
    lib = mb.load_library()
    rc = lib.mnemcu_trans_close_w(None, 3, 1)
    assert rc == 1 and b"NULL" in lib.mnemcu_last_error()
    x = np.zeros(40 * 40, dtype=np.uint8)
    rc = lib.mnemcu_trans_close_w(_lib.u8p(x), 40, 1)          # S > 32 is outside this build
    assert rc == 1


def test_mod_data_natural_order():
    pert, S, labs = mb.mod_data(["g10", "g2", "g1", "g2", "g10"])
    assert labs == ["g1", "g2", "g10"] and S == 3
    assert pert.tolist() == [3, 2, 1, 2, 3]
    with pytest.raises(ValueError):
        mb.mod_data(["1_2", "1"])                      # one integer per cell cannot express a double knock-down


def test_get_rho_follows_getRho():
    """getRho() R/mnems_low.r:308-324 on "_"-joined labels; S-genes in natural order."""
    Rho, sg = mb.get_rho(["2", "10_2", "1", "10", "1_2_10"])
    assert sg == ["1", "2", "10"]
    assert Rho.tolist() == [[0, 0, 1, 0, 1], [1, 1, 0, 0, 1], [0, 1, 0, 1, 1]]


# ---- a compiled C++ consumer of the ABI (tests/cabi/cabi_check.cpp, built by __graft_entry__.build()) ---------------
CABI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cabi", "cabi_check")


def _cabi(*args):
    import subprocess
    if not os.path.exists(CABI):
        import __graft_entry__ as ge
        ge.build()
    return subprocess.run([CABI] + list(args), capture_output=True, text=True, timeout=300)


def test_struct_layouts_agree_between_c_and_ctypes():
    """sizeof / offsetof as the C++ compiler sees include/mnemcu.h == what the ctypes binding declares."""
    import ctypes as C
    from mnem_b200 import _lib
    out = dict(line.split(None, 1) for line in _cabi("layout").stdout.strip().splitlines())
    o = [int(x) for x in out["em_opts"].split()]
    f = _lib.EmOpts
    assert o == [C.sizeof(f), f.complete.offset, f.logtype.offset, f.batch.offset, f.max_trace.offset, f.max_iter_guard.offset,
                 f.next_start.offset, f.next_start_user.offset, f.affinity.offset, f.tape.offset, f.tape_cap.offset, f.probs_mode.offset, f.mw.offset, f.nullcomp.offset, f.theta.offset]
    s = [int(x) for x in out["em_stats"].split()]
    g = _lib.EmStats
    assert s == [C.sizeof(g), g.em_iters.offset, g.ms_total.offset, g.estep_bytes.offset, g.tape_words.offset]
    assert int(out["version"]) == mb.load_library().mnemcu_version()


def test_cpp_consumer_sees_clean_errors_without_a_device():
    """No CPU fallback: from C++ too, every compute entry fails with MNEMCU_ERR_CUDA and a message; argument errors come
    first.  (On a GPU box the program reports the device instead.)"""
    r = _cabi("nogpu")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "has_gpu" in r.stdout or ("ctx_create rc=2" in r.stdout and "get_ll rc=2" in r.stdout and "null_pert rc=1" in r.stdout)


def test_rcpp_shim_compiles_against_the_stub():
    """rshim/mnemcu_shim.cpp builds (g++ -Wall -Wextra) against tests/cabi/rcpp_stub/Rcpp.h and links libmnemcu.so."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "cabi", "shim_check")
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build()
    r = subprocess.run([exe, "syntax"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
