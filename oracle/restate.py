"""oracle/restate.py -- TEST INFRASTRUCTURE: ctypes wrapper of oracle/liblu_oracle.so (the plain-C restatement
in lu_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
import ctypes
import os
import subprocess
import numpy as np
from . import layout

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liblu_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "lu_oracle.c")
    if force or not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liblu_oracle.so"])
    return _PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_PATH)
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def init_matrix(N, v, Px=1, Py=1, Pz=1, seed=42):
    N, v, Px, Py, Pz = int(N), int(v), int(Px), int(Py), int(Pz)
    d = layout.dims(N, v, Px, Py, Pz)
    A = np.zeros((d["P"], d["Ml"] * d["Nl"]))
    lib().oracle_init_matrix(N, v, Px, Py, Pz, seed, _dp(A))
    return [A[r].reshape(d["Ml"], d["Nl"]) for r in range(d["P"])]


def lu(A_locals, N, v, Px=1, Py=1, Pz=1):
    """A_locals: list of P (Ml x Nl) arrays in rank order.  Returns dict(C=[...], perm)."""
    N, v, Px, Py, Pz = int(N), int(v), int(Px), int(Py), int(Pz)
    d = layout.dims(N, v, Px, Py, Pz)
    P, loc = d["P"], d["Ml"] * d["Nl"]
    A = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float64).reshape(-1) for a in A_locals]))
    C = np.zeros((P, loc))
    perm = np.full(d["M"], -1, dtype=np.int32)
    rc = lib().oracle_lu(N, v, Px, Py, Pz, _dp(A), _dp(C), _ip(perm))
    if rc != 0:
        raise ValueError("oracle_lu: unsupported grid/tile (needs Px == Py and v % Pz == 0)")
    return dict(C=[C[r].reshape(d["Ml"], d["Nl"]) for r in range(P)], perm=perm, dims=d)


# ---- building blocks (for the golden-vector tests) ------------------------------------------------
def push_pivots_up(mat, cur_pivots, fnpr):
    m = np.ascontiguousarray(np.asarray(mat, dtype=np.float64).copy())
    cp = np.ascontiguousarray(np.asarray(cur_pivots, dtype=np.int32))
    lib().oracle_push_pivots_up(_dp(m), m.shape[0], m.shape[1], _ip(cp), int(fnpr))
    return m


def inverse_permute_rows(mat, perm, out_rows, out_cols):
    m = np.ascontiguousarray(np.asarray(mat, dtype=np.float64))
    p = np.ascontiguousarray(np.asarray(perm, dtype=np.int32))
    out = np.zeros((out_rows, out_cols))
    lib().oracle_inverse_permute_rows(_dp(m), _dp(out), m.shape[1], out_rows, out_cols, _ip(p))
    return out


def permute_rows(mat, perm, out_rows, out_cols):
    m = np.ascontiguousarray(np.asarray(mat, dtype=np.float64))
    p = np.ascontiguousarray(np.asarray(perm, dtype=np.int32))
    out = np.zeros((out_rows, out_cols))
    lib().oracle_permute_rows(_dp(m), _dp(out), m.shape[0], m.shape[1], out_cols, _ip(p))
    return out


def butterfly_pair(pi, r, Px):
    return lib().oracle_butterfly_pair(pi, r, Px)


def g2l_owner(grows, Px, v):
    g = np.ascontiguousarray(np.asarray(grows, dtype=np.int32))
    own = np.zeros(len(g), dtype=np.int32)
    lib().oracle_g2l_owner(_ip(g), len(g), Px, v, _ip(own))
    return own


def getrf_perm(cand, n, v):
    """cand: n x (v+1) (column 0 = tags).  Returns (perm[max(2v,n)], factored n x v)."""
    a = np.ascontiguousarray(np.asarray(cand, dtype=np.float64)[:n, 1:v + 1].copy())
    if a.size == 0:
        a = np.zeros((1, v))
    perm = np.zeros(max(2 * v, n), dtype=np.int32)
    lib().oracle_getrf_perm(n, v, _dp(a), v, _ip(perm))
    return perm, a
