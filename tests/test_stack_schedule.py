"""CPU: the column-owner schedule of stack_getrf_kernel (oracle/stack_ref.py) selects the pivots LAPACK's dgetrf selects
(oracle/restate.getrf_perm = the reference's LUP, conflux_opt.hpp:143-166), including integer matrices full of ties,
and leaves L\\U in place."""
import numpy as np
import pytest

from oracle import restate, stack_ref


@pytest.mark.parametrize("n,v,kind", [(64, 32, "normal"), (96, 48, "normal"), (128, 64, "ties"), (64, 64, "ties"),
                                      (200, 32, "ties"), (160, 80, "zeros")])
def test_column_owner_schedule_matches_dgetrf(n, v, kind):
    rng = np.random.default_rng(n + v)
    if kind == "normal":
        P = rng.standard_normal((n, v))
    elif kind == "ties":
        P = rng.integers(-2, 3, size=(n, v)).astype(np.float64)
    else:                                          # padded stacks: half of the rows are zero
        P = rng.standard_normal((n, v))
        P[rng.permutation(n)[: n // 2]] = 0.0
    perm, W = stack_ref.stack_getrf(P)
    cand = np.concatenate([np.zeros((n, 1)), P], axis=1)
    perm_ref, lu_ref = restate.getrf_perm(cand, n, v)
    assert np.array_equal(perm, perm_ref[:v])
    scale = max(1.0, np.abs(P).max())
    assert np.abs(W[perm] - lu_ref[:v]).max() <= 1e-11 * scale * v      # L00\U00 of the winners
