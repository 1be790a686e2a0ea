"""oracle/ozaki_ref.py -- TEST INFRASTRUCTURE: exact numpy restatement of the int8 digit-plane trailing update of
conflux_b200/csrc/ozaki.cu (C -= L*U through 8 signed digit planes per operand, 36 exact integer plane products grouped
by weight, FP64 recombination in the kernel's operation order).  Everything up to the recombination is integer-exact and
every FP64 operation of the kernel is reproduced in the same order, so the CUDA result must match BIT FOR BIT.
(The reference itself calls cblas_dgemm here, conflux_opt.hpp:1628-1632; agreement with plain FP64 is checked separately
within a rounding-level tolerance.)"""
import numpy as np

S = 8


def split(X):
    """X[o][k] -> planes int8 [S][o][k], exps int32 [o]:  X = 2^(e-6) * sum_s d_s 2^(-7s)  (to 55 bits of the row max)."""
    X = np.asarray(X, dtype=np.float64)
    mx = np.max(np.abs(X), axis=1) if X.shape[1] else np.zeros(X.shape[0])
    _, e = np.frexp(mx)
    e = np.where(mx > 0, e, 0).astype(np.int32)
    r = X * np.exp2(-e.astype(np.float64))[:, None] * 64.0
    planes = np.zeros((S,) + X.shape, dtype=np.int8)
    for s in range(S):
        d = np.rint(r)
        planes[s] = d.astype(np.int8)
        r = (r - d) * 128.0
    return planes, e


def gemm(AT, B, C):
    """D = C - AT^T @ B exactly as the kernel computes it."""
    pa, ea = split(np.ascontiguousarray(AT.T))
    pb, eb = split(np.ascontiguousarray(B.T))
    M, N = AT.shape[1], B.shape[1]
    total = np.zeros((M, N))
    w = 1.0
    for g in range(S):
        acc = np.zeros((M, N), dtype=np.int64)
        for s in range(g + 1):
            acc += pa[s].astype(np.int64) @ pb[g - s].astype(np.int64).T
        assert np.abs(acc).max(initial=0) < 2 ** 31
        total = acc.astype(np.float64) * w + total          # product exact (power of two): one rounding, like the fma
        w *= 0.0078125
    patch = total * np.exp2(ea.astype(np.float64) - 12.0)[:, None]       # exact scaling
    return C - patch * np.exp2(eb.astype(np.float64))[None, :], (pa, pb, ea, eb)
