"""tests/golden/make_golden_bench.py -- pivot sequences of THE REFERENCE ITSELF (oracle/_ref/libref_lu.so = the
/root/reference sources compiled by oracle/build_ref.sh, ranks = threads) at the BASELINE bench configurations that
fit this container's memory:
    C2  N=16384 v=256 grid 1x1x1   (bench.py --gpus 1)
    C3  N=32768 v=512 grid 2x2x1   (bench.py --gpus 4)
Written to tests/golden/lu_perms_bench.npz (int32 permutations only, ~200 KB).  bench.py's parity block and
tests/test_gpu_lu.py compare the GPU pivot sequence with these element by element.  The 2- and 8-GPU configurations
(1x1x2 at N=32768: 80 GiB of host buffers; 2x2x2 at N=65536) do not fit here and stay residual-checked only.
Run in the build container only (needs /root/reference):  python tests/golden/make_golden_bench.py [C2] [C3]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "lu_perms_bench.npz")
CASES = {"C2": (16384, 256, 1, 1, 1), "C3": (32768, 512, 2, 2, 1)}

if __name__ == "__main__":
    if not ref.available():
        sys.exit("oracle/_ref/libref_lu.so missing: run oracle/build_ref.sh first")
    want = [a for a in sys.argv[1:] if a in CASES] or list(CASES)
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for name in want:
        N, v, Px, Py, Pz = CASES[name]
        t0 = time.time()
        r = ref.lu_run(N, v, Px, Py, Pz, want_factors=False, blas_threads=max(1, (os.cpu_count() or 1) // (Px * Py * Pz)))
        perm = r["perm"]
        assert sorted(perm.tolist()) == list(range(r["dims"]["M"])), name
        out[name] = perm.astype(np.int32)
        out[name + "_case"] = np.array([N, v, Px, Py, Pz], dtype=np.int32)
        np.savez_compressed(OUT, **out)
        print(f"{name}: N={N} v={v} grid {Px}x{Py}x{Pz}  reference LU_rep {r['ms']:.0f} ms (wall {time.time() - t0:.0f} s)", flush=True)
