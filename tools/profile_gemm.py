"""tools/profile_gemm.py -- the trailing-update GEMM alone at the first-step shape of the N=1 bench workload, for
`ncu --set full -k regex:gemm_tn -s 1 -c 1` (launch 0 is the warm-up)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import conflux_b200 as cb
M = N = int(sys.argv[1]) if len(sys.argv) > 1 else 16128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.default_rng(0)
_, ms = cb.dbg.gemm_tn(rng.standard_normal((K, M)), rng.standard_normal((K, N)), rng.standard_normal((M, N)), -1.0, 1.0, reps=1)
print(f"gemm_tn {M}x{N}x{K}: {ms:.3f} ms (under the profiler: not a benchmark number)")
