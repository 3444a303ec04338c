"""mnem_b200 -- B200-native (sm_100a) implementation of M&NEM's EM hot path behind the reference's entry points.

The package holds only what the path needs: csrc/ (CUDA kernels + the C ABI of include/mnemcu.h, built into
libmnemcu.so), the ctypes binding (_lib.py), the host mirror of the reference's R functions (api.py) and the
simData-style input generator (simdata.py).  There is no CPU fallback.
"""
from ._lib import MnemcuError, load_library  # noqa: F401
from .api import (Context, bootstrap, device_count, doMean, eigenMapMatMult, em_run, getAffinity, getIC, getLL, getProbs,  # noqa: F401
                  get_rho, get_sgenes, launch_count, maxCol_row, mnem, mnemk, mod_data, neighbours, nem, nem_greedy, sample_init, scoreAdj,
                  transClose_Del, transClose_Ins, transClose_W, transitive_closure, transitive_reduction)
