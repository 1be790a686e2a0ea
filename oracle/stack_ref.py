"""oracle/stack_ref.py -- TEST INFRASTRUCTURE (never imported by the product).

numpy restatement of the SCHEDULE of stack_getrf_kernel (conflux_b200/csrc/panel.cu): the 2v x v tournament stack of
/root/reference/src/conflux/lu/conflux_opt.hpp:291 (LUP = LAPACKE_dgetrf + ipiv -> perm, :143-166) factored by COLUMN-block
owners.  Rows never move; LAPACK's interchanges live in a per-row "position" that every owner replays from the published
(pivot row, position) pairs, so idamax tie-breaking (first maximal |a| in swapped order) is reproduced.  Used by
tests/test_stack_schedule.py to pin the schedule against oracle/restate.getrf_perm on CPU; the CUDA kernel itself is
compared with the row-owner kernel and the oracle in tests/test_gpu_kernels.py.
"""
import numpy as np

CB = 16  # columns per owner (SK_CB)


def stack_getrf(P):
    """P: (n, v) float64, n >= v, v % 16 == 0.  Returns (perm[v], W) with W the in-place L\\U in unpermuted row order."""
    P = np.array(P, dtype=np.float64)
    n, v = P.shape
    assert n >= v and v % CB == 0
    W = P.copy()
    perm = np.zeros(v, dtype=np.int64)
    ppos = np.zeros(v, dtype=np.int64)
    for c in range(v // CB):                       # one owner per block column; owner c only reads blocks b < c
        c0 = c * CB
        a = W[:, c0:c0 + CB].copy()                # the owner's registers: its 16 columns of every row
        pos = np.arange(n)
        mypiv = np.full(n, -1)
        for b in range(c):                         # apply the published blocks in order
            pivs, pp = perm[b * CB:(b + 1) * CB], ppos[b * CB:(b + 1) * CB]
            L = W[:, b * CB:(b + 1) * CB]          # published multipliers (pivot rows: their L\U row)
            for k in range(CB):                    # replay the interchanges of block b
                live = (mypiv < 0)
                live[pivs[:k + 1]] = False
                hit = live & (pos == b * CB + k)
                pos[hit] = pp[k]
            U12 = a[pivs].copy()                   # forward substitution with the publisher's unit-lower L11
            L11 = L[pivs]
            for i in range(1, CB):
                for s in range(i):
                    U12[i] -= L11[i, s] * U12[s]
            upd = (mypiv < 0)
            upd[pivs] = False
            a[upd] -= L[upd] @ U12                 # (the kernel: 16 fma per element, k ascending)
            a[pivs] = U12
            mypiv[pivs] = b * CB + np.arange(CB)
        for j in range(CB):                        # the owner's own block: partial pivoting inside one CTA
            jg = c0 + j
            cand = np.where(mypiv < 0)[0]
            key = np.abs(a[cand, j])
            best = cand[(key == key.max())]
            win = best[np.argmin(pos[best])]       # first maximal |a| in LAPACK's swapped order
            perm[jg], ppos[jg] = win, pos[win]
            mypiv[win] = jg
            live = (mypiv < 0)
            moved = live & (pos == jg)
            pos[moved] = ppos[jg]
            piv = a[win, j]
            if piv != 0.0:
                a[live, j] = a[live, j] * (1.0 / piv)
                a[np.ix_(live, range(j + 1, CB))] -= np.outer(a[live, j], a[win, j + 1:])
        W[:, c0:c0 + CB] = a
    return perm, W
