"""tools/stack_speed.py -- the 2v x v tournament stack on both pivot-search kernels (isolated, one GPU)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import conflux_b200 as cb

rng = np.random.default_rng(1)
for (n, v) in [(1024, 512), (512, 256), (256, 128)]:
    P = rng.standard_normal((n, v))
    out = {}
    for mode in ("1", "0"):
        os.environ["CFLX_STACK_KERNEL"] = mode
        perm, A00, LU, ms = cb.dbg.panel(P, reps=20)
        out[mode] = (perm, A00, ms)
    same = np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1])
    print(f"stack {n}x{v}: column-owner kernel {out['1'][2]*1e3:.0f} us, row-owner kernel {out['0'][2]*1e3:.0f} us, "
          f"pivots and L00\\U00 bit-identical: {same}")
