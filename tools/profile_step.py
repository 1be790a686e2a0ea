"""tools/profile_step.py -- one LU factorisation of the N=1 bench workload, for `ncu` (launch list / full capture)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import conflux_b200 as cb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
v = int(sys.argv[2]) if len(sys.argv) > 2 else 256
comm = cb.Comm(1, 0, None, 0)
gv = cb.lu_params(N, N, v, 1, 1, 1, comm)
perm = np.zeros(gv.M, dtype=np.int32)
ms = cb.LU_rep(gv, None, perm)
print(f"N={N} v={v}: {ms:.2f} ms (under the profiler: not a benchmark number)")
