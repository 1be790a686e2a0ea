"""tools/ozaki_speed.py -- one trailing update of the first-step shape of a BASELINE config on both GEMM paths (for ncu):
    python tools/ozaki_speed.py [M] [N] [K] [reps]"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import conflux_b200 as cb

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16128
K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rng = np.random.default_rng(1)
AT = rng.uniform(-1, 1, (K, M))
B = rng.standard_normal((K, N))
r = cb.dbg.ozaki_gemm(AT, B, None, reps=reps)
print(f"ozaki {M}x{N}x{K}: {r['ms']:.3f} ms = {2.0 * M * N * K / (r['ms'] * 1e-3) / 1e12:.1f} TFLOP/s FP64-equivalent "
      f"({36 * 2.0 * M * N * K / (r['ms'] * 1e-3) / 1e15:.2f} POP/s int8), digit planes {r['split_ms']:.3f} ms")
_, dm = cb.dbg.gemm_tn(AT, B, None, -1.0, 0.0, reps=reps)
print(f"dmma  {M}x{N}x{K}: {dm:.3f} ms = {2.0 * M * N * K / (dm * 1e-3) / 1e12:.1f} TFLOP/s")
