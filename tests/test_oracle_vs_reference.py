"""CPU, build container only: the restatement oracle against the LIVE reference (oracle/_ref/libref_lu.so built
from /root/reference by oracle/build_ref.sh).  Skipped where the prebuilt library is absent."""
import numpy as np
import pytest

from oracle import layout, ref, restate

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")

CASES = [(48, 4, 1, 1, 1), (128, 16, 2, 2, 1), (128, 8, 2, 2, 2), (96, 16, 1, 1, 2), (72, 4, 3, 3, 2), (256, 32, 4, 4, 1)]


@pytest.mark.parametrize("case", CASES)
def test_matches_live_reference(case):
    N, v, Px, Py, Pz = case
    r = ref.lu_run(N, v, Px, Py, Pz)
    A0 = restate.init_matrix(N, v, Px, Py, Pz)
    for a, b in zip(A0, r["A"]):
        assert np.array_equal(a, b)                          # generator is bit-identical to lu_params::InitMatrix
    o = restate.lu(r["A"], N, v, Px, Py, Pz)
    assert np.array_equal(o["perm"], r["perm"])
    for a, b in zip(o["C"], r["C"]):
        assert np.abs(a - b).max() <= 1e-10 * 6.0
    A = layout.assemble(r["A"], N, v, Px, Py, Pz)
    assert layout.residual(A, layout.assemble(o["C"], N, v, Px, Py, Pz), o["perm"]) <= 1e-12
