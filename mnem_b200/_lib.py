"""ctypes binding of libmnemcu.so -- exactly the symbols declared in include/mnemcu.h.

This is the stand-in for the Rcpp shim (rshim/mnemcu_shim.cpp): R is not installable in this image, so the
reference-facing host layer is exercised from Python instead.  There is no fallback: if the library cannot be
loaded, or a call fails, MnemcuError is raised with the library's message.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MNEMCU_LIB: another build of the same library (A/B timing of kernel changes); never a different implementation
LIB_PATH = os.environ.get("MNEMCU_LIB") or os.path.join(_HERE, "libmnemcu.so")

c_dp = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_i8p = C.POINTER(C.c_int8)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_ctx = C.c_void_p


class MnemcuError(RuntimeError):
    pass


NEXT_START_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p)


class EmOpts(C.Structure):
    _fields_ = [("complete", C.c_int), ("logtype", C.c_double), ("batch", C.c_int), ("max_trace", C.c_int),
                ("max_iter_guard", C.c_int), ("next_start", NEXT_START_FN), ("next_start_user", C.c_void_p),
                ("affinity", C.c_int), ("tape", C.POINTER(C.c_int32)), ("tape_cap", C.c_int64),
                ("probs_mode", C.c_int), ("mw", C.POINTER(C.c_double)), ("nullcomp", C.c_int), ("theta", C.POINTER(C.c_int32))]


class EmStats(C.Structure):
    _fields_ = [("em_iters", C.c_int64), ("cand_scored", C.c_int64), ("greedy_iters", C.c_int64),
                ("passes_estep", C.c_int64), ("passes_mstats", C.c_int64), ("ms_total", C.c_double),
                ("ms_estep", C.c_double), ("ms_mstats", C.c_double), ("ms_search", C.c_double),
                ("ms_mixture", C.c_double), ("estep_bytes", C.c_double), ("search_sm_busy", C.c_double),
                ("tape_words", C.c_int64)]


# name -> (restype, argtypes); every symbol of include/mnemcu.h
SIGNATURES = {
    "mnemcu_version": (C.c_int, []),
    "mnemcu_last_error": (C.c_char_p, []),
    "mnemcu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mnemcu_launch_count": (C.c_int64, []),
    "mnemcu_trans_close_w": (C.c_int, [c_u8p, C.c_int, C.c_int]),
    "mnemcu_trans_close_ins": (C.c_int, [c_u8p, C.c_int, C.c_int, c_i32p, c_i32p]),
    "mnemcu_trans_close_del": (C.c_int, [c_u8p, C.c_int, C.c_int, c_i32p, c_i32p]),
    "mnemcu_max_col_row": (C.c_int, [c_dp, C.c_int, C.c_int, c_i32p]),
    "mnemcu_matmul": (C.c_int, [c_dp, c_dp, C.c_int, C.c_int, C.c_int, c_dp]),
    "mnemcu_trans_reduce": (C.c_int, [c_u8p, C.c_int, C.c_int, c_u8p]),
    "mnemcu_neighbours": (C.c_int, [c_u8p, C.c_int, c_u8p, c_i8p, c_i32p]),
    "mnemcu_get_affinity": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, C.c_int, C.c_int, C.c_double, c_dp]),
    "mnemcu_get_ll": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, C.c_int, C.c_double, c_dp]),
    "mnemcu_score_adj": (C.c_int, [c_dp, C.c_int, C.c_int, c_u8p, C.c_int, C.c_int, c_dp, c_i32p, c_dp]),
    "mnemcu_score_adj_marginal": (C.c_int, [c_dp, C.c_int, C.c_int, c_u8p, C.c_int, C.c_int, C.c_double, c_dp, c_i32p, c_dp]),
    "mnemcu_nem_greedy_marginal": (C.c_int, [c_dp, C.c_int, C.c_int, c_u8p, C.c_int, C.c_double, c_u8p, c_i32p, c_dp, c_dp,
                                             c_i32p, c_i64p]),
    "mnemcu_nem_greedy": (C.c_int, [c_dp, C.c_int, C.c_int, c_u8p, C.c_int, c_u8p, c_i32p, c_dp, c_dp, c_i32p, c_i64p]),
    "mnemcu_ctx_create": (C.c_int, [c_dp, C.c_int, C.c_int, c_i32p, C.c_int, C.c_int, C.POINTER(c_ctx)]),
    "mnemcu_ctx_create_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_i32p, C.c_int, C.c_int, C.POINTER(c_ctx)]),
    "mnemcu_ctx_create_rho": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, C.c_int, C.c_int, C.POINTER(c_ctx)]),
    "mnemcu_ctx_create_synthetic": (C.c_int, [C.c_int, C.c_int, c_i32p, C.c_int, c_u8p, C.c_int, c_u8p, c_i32p,
                                              C.c_uint64, C.c_int, C.POINTER(c_ctx)]),
    "mnemcu_synthetic_fill_host": (C.c_int, [c_dp, C.c_int, C.c_int64, C.c_int64, c_i32p, C.c_int, c_u8p, C.c_int,
                                             c_u8p, c_i32p, C.c_uint64, C.c_int]),
    "mnemcu_ctx_destroy": (C.c_int, [c_ctx]),
    "mnemcu_ctx_dims": (C.c_int, [c_ctx, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mnemcu_ctx_read_data": (C.c_int, [c_ctx, C.c_int64, C.c_int64, c_dp]),
    "mnemcu_do_mean": (C.c_int, [c_ctx, c_dp, C.c_int, c_dp]),
    "mnemcu_get_probs": (C.c_int, [c_ctx, C.c_int, C.c_int, c_u8p, c_i32p, c_dp, c_dp, C.c_int, C.c_double, c_dp, c_dp,
                                   c_dp, c_i32p]),
    "mnemcu_em_run": (C.c_int, [c_ctx, C.c_int, C.c_int, c_u8p, C.POINTER(EmOpts), c_u8p, c_i32p, c_dp, c_dp, c_dp,
                                c_i32p, c_dp, c_dp, c_dp, c_dp, c_i32p, C.POINTER(EmStats)]),
    "mnemcu_bench_estep": (C.c_int, [c_ctx, C.c_int, C.c_int, c_u8p, c_i32p, C.c_int, c_dp, c_dp]),
    "mnemcu_bench_mstats": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, c_dp, c_dp]),
}

_lib = None


def load_library():
    """Load libmnemcu.so (built by __graft_entry__.build()).  Raises if it is missing: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MnemcuError("libmnemcu.so is not built (%s); run `python __graft_entry__.py`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load_library().mnemcu_last_error()
        raise MnemcuError("libmnemcu error %d: %s" % (rc, msg.decode("utf-8", "replace") if msg else "?"))


def dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def u8p(a):
    return a.ctypes.data_as(c_u8p) if a is not None else None


def i32p(a):
    return a.ctypes.data_as(c_i32p) if a is not None else None
