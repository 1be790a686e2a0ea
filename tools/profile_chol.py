"""tools/profile_chol.py -- one Cholesky factorisation (N, v) on one GPU, for `ncu` launch lists."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conflux_b200 as cb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
v = int(sys.argv[2]) if len(sys.argv) > 2 else 512
comm = cb.Comm(1, 0, None, 0)
ch = cb.cholesky.initialize(N, v, (1, 1, 1), comm)
ms = ch.parallelCholesky()
ms2 = ch.parallelCholesky(upload=False)
print(f"cholesky N={ch.N} v={ch.v}: {ms:.2f} / {ms2:.2f} ms (under a profiler this is not a benchmark number)")
ch.finalize()
comm.close()
