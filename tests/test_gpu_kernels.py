"""GPU: every hand-written kernel in isolation, through the C ABI, against numpy / the restatement oracle."""
import numpy as np
import pytest

import conflux_b200 as cb
from oracle import restate

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (64, 32, 4), (200, 130, 20), (257, 384, 64), (1000, 1024, 256),
                                   (8, 8, 8), (130, 2, 12), (1, 256, 128), (513, 514, 512)])
def test_gemm_tn_matches_numpy(M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    AT, B, C = rng.standard_normal((K, M)), rng.standard_normal((K, N)), rng.standard_normal((M, N))
    for alpha, beta in [(-1.0, 1.0), (1.0, 0.0), (0.5, -2.0)]:
        D, _ = cb.dbg.gemm_tn(AT, B, C, alpha, beta)
        ref = beta * C + alpha * (AT.T @ B)
        assert np.abs(D - ref).max() <= 1e-13 * K * 16, (M, N, K, alpha, beta)


def _check_panel(P, tol=1e-11):
    n, v = P.shape
    perm, A00, LU, _ = cb.dbg.panel(P)
    cand = np.concatenate([np.zeros((n, 1)), P], axis=1)
    perm_ref, lu_ref = restate.getrf_perm(cand, n, v)
    assert np.array_equal(perm, perm_ref[:v]), (n, v, perm[:8], perm_ref[:8])
    if n >= v:
        # oracle's factored panel has rows in pivoted order; top v x v is L00\U00
        scale = max(1.0, np.abs(P).max())
        assert np.abs(A00 - lu_ref[:v]).max() <= tol * scale * v
        # rows never move on the GPU: LU[perm[i]] is pivoted row i; the others hold their multipliers
        assert np.abs(LU[perm][:, :] - lu_ref[:v])[np.tril_indices(v, -1)].max(initial=0) <= tol * scale * v


@pytest.mark.parametrize("n,v", [(8, 4), (64, 8), (200, 16), (300, 32), (1024, 32), (4096, 64), (20000, 128),
                                 (1024, 512), (33000, 64)])
def test_panel_getrf_random(n, v):
    rng = np.random.default_rng(n + v)
    _check_panel(5.0 + rng.random((n, v)))


def test_panel_getrf_integer_ties():
    rng = np.random.default_rng(3)
    for (n, v) in [(16, 4), (64, 8), (300, 16), (700, 32)]:
        _check_panel(rng.integers(0, 4, size=(n, v)).astype(np.float64) + np.eye(n, v) * 0.0)


@pytest.mark.parametrize("n,v", [(1024, 512), (512, 256), (256, 128), (700, 512), (1000, 16), (96, 32), (512, 512)])
def test_column_owner_kernel_is_bit_identical_to_the_row_owner_kernel(n, v, monkeypatch):
    """Panels of <= 1024 rows (the 2v x v tournament stacks) run on stack_getrf_kernel (one CTA per 16 columns, no
    exchange per column); its fma chains are those of panel_getrf_kernel, so pivots AND factors must agree bit for bit."""
    rng = np.random.default_rng(n * 7 + v)
    for P in (rng.standard_normal((n, v)), rng.integers(-3, 4, size=(n, v)).astype(np.float64)):
        monkeypatch.setenv("CFLX_STACK_KERNEL", "1")
        perm1, A1, LU1, _ = cb.dbg.panel(P)
        monkeypatch.setenv("CFLX_STACK_KERNEL", "0")
        perm0, A0, LU0, _ = cb.dbg.panel(P)
        assert np.array_equal(perm1, perm0)
        assert np.array_equal(A1, A0)                      # L00\\U00 of the winners, every bit
        # the factored panel: multipliers of every row (the row-owner kernel never writes the U part of a pivot row back)
        rest = np.setdiff1d(np.arange(n), perm1)
        assert np.array_equal(LU1[rest], LU0[rest])
        low = np.tril_indices(v, -1)
        assert np.array_equal(LU1[perm1][low], LU0[perm0][low])
        assert np.array_equal(np.triu(LU1[perm1]), np.triu(A1))   # and the column-owner kernel leaves L\\U in place


def test_panel_getrf_fewer_rows_than_columns():
    rng = np.random.default_rng(5)
    for (n, v) in [(3, 8), (1, 4), (10, 32), (0, 8)]:
        P = rng.standard_normal((n, v))
        perm, _, _, _ = cb.dbg.panel(P)
        cand = np.concatenate([np.zeros((n, 1)), P], axis=1)
        perm_ref, _ = restate.getrf_perm(cand, n, v)
        assert np.array_equal(perm, perm_ref[:v]), (n, v)


@pytest.mark.parametrize("n,v", [(16, 4), (100, 8), (257, 32), (1000, 64), (3000, 128), (2048, 256), (777, 512)])
def test_trsm_matches_numpy(n, v):
    rng = np.random.default_rng(n + 31 * v)
    L = np.tril(rng.uniform(-1, 1, (v, v)) * min(1.0, 8.0 / v), -1) + np.eye(v)   # keep cond(L) modest
    U = np.triu(rng.uniform(-1, 1, (v, v))) + np.diag(np.sign(rng.standard_normal(v)) * (2 + rng.random(v)))
    A00 = np.tril(L, -1) + U
    B, R = rng.standard_normal((n, v)), rng.standard_normal((v, n))
    X, Y = cb.dbg.trsm(A00, B, R)
    assert np.abs(X @ U - B).max() <= 1e-10 * np.abs(B).max() * v
    assert np.abs(L @ Y - R).max() <= 1e-10 * np.abs(R).max() * v


def test_push_pivots_reference_unit_vector_and_seeded_cases(golden_dir):
    """plan_moves + push_phase1..3 in isolation on the reference's own unit-test input (tests/unit/test_utils.cpp:8-84:
    pivots {2,1,5}, expected row order [2,1,5,3,4,0,6,7]) and on the seeded cases the reference itself produced
    (tests/golden/helpers.npz; the kernels move 128-bit pairs, so the odd-width cases are run with one padding column)."""
    import os
    H = np.load(os.path.join(golden_dir, "helpers.npz"))
    A, gri, a01 = cb.dbg.push_pivots(H["push0_in"], H["push0_piv"][1:], 0)
    assert np.array_equal(A, H["push0_in"][[2, 1, 5, 3, 4, 0, 6, 7]]) and np.array_equal(A, H["push0_out"])
    assert gri.tolist() == [2, 1, 5, 3, 4, 0, 6, 7]
    assert np.array_equal(a01, H["push0_in"][[2, 1, 5]])
    for i in range(int(H["push_n"])):
        m, cp, fnpr, out = H[f"push{i}_in"], H[f"push{i}_piv"], int(H[f"push{i}_fnpr"]), H[f"push{i}_out"]
        pad = m.shape[1] & 1
        mp = np.concatenate([m, np.arange(m.shape[0], dtype=np.float64)[:, None]], axis=1) if pad else m
        A, gri, _ = cb.dbg.push_pivots(mp, cp[1:1 + cp[0]], fnpr)
        assert np.array_equal(A[:, :m.shape[1]], out), i
        if pad:
            assert np.array_equal(A[:, -1], gri.astype(np.float64)), i      # the padding column travelled with its row


def test_fp64_pipe_probe_runs():
    dmma, dfma = cb.dbg.fp64_peak(0), cb.dbg.fp64_peak(1)
    print(f"FP64 peaks: DMMA {dmma:.1f} TFLOP/s, DFMA {dfma:.1f} TFLOP/s")
    assert dmma > 1.0 and dfma > 1.0
