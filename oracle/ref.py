"""oracle/ref.py -- TEST INFRASTRUCTURE: ctypes wrapper of oracle/_ref/libref_lu.so, i.e. the reference's own
LU_rep / lu_params compiled by oracle/build_ref.sh (ranks = threads).  Only tests/, tests/golden/make_golden.py,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
import ctypes
import os
import numpy as np
from . import layout

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_lu.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def lu_run(N, v, Px=1, Py=1, Pz=1, n_rep=1, blas_threads=1, want_factors=True):
    """Returns dict(A=[P local arrays], C=[P local arrays], perm, ms, dims)."""
    N, v, Px, Py, Pz = int(N), int(v), int(Px), int(Py), int(Pz)
    d = layout.dims(N, v, Px, Py, Pz)
    P, loc = d["P"], d["Ml"] * d["Nl"]
    A = np.zeros((P, loc)) if want_factors else None
    C = np.zeros((P, loc)) if want_factors else None
    perm = np.full(d["M"], -1, dtype=np.int32)
    ms = ctypes.c_double(0)
    lib().ref_lu_run(N, v, Px, Py, Pz, n_rep, _dp(A) if want_factors else None, _dp(C) if want_factors else None,
                     _ip(perm), ctypes.byref(ms), blas_threads)
    out = dict(perm=perm, ms=ms.value, dims=d)
    if want_factors:
        out["A"] = [A[r].reshape(d["Ml"], d["Nl"]) for r in range(P)]
        out["C"] = [C[r].reshape(d["Ml"], d["Nl"]) for r in range(P)]
    return out


def init_matrix(N, v, Px=1, Py=1, Pz=1):
    d = layout.dims(N, v, Px, Py, Pz)
    A = np.zeros((d["P"], d["Ml"] * d["Nl"]))
    lib().ref_init_matrix(N, v, Px, Py, Pz, _dp(A))
    return [A[r].reshape(d["Ml"], d["Nl"]) for r in range(d["P"])]


def lu_bench(N, v, Px=1, Py=1, Pz=1, n_warm=1, n_rep=3, budget_s=300.0, blas_threads=1):
    """Time-boxed repetition loop inside ONE lu_params object (bench.py --impl reference).  Returns dict(inner_ms=[...],
    outer_ms=[...]) of the timed repetitions actually done: inner = what LU_rep returns (conflux_opt.hpp:1807)."""
    inner = np.zeros(max(1, n_rep))
    outer = np.zeros(max(1, n_rep))
    done = ctypes.c_int(0)
    lib().ref_lu_bench.argtypes = [ctypes.c_int] * 7 + [ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                                                        ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib().ref_lu_bench(int(N), int(v), int(Px), int(Py), int(Pz), int(n_warm), int(n_rep), float(budget_s), _dp(inner),
                       _dp(outer), ctypes.byref(done), int(blas_threads))
    return dict(inner_ms=inner[:done.value].tolist(), outer_ms=outer[:done.value].tolist())
