"""tools/ozaki_probe.py -- GPU-side bring-up probe of the int8 tcgen05 trailing update: smallest case first, prints
where the CUDA result departs from the exact restatement (planes, exponents, tiles)."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import conflux_b200 as cb
from oracle import ozaki_ref

for (M, N, K, kind) in [(128, 64, 128, "ints"), (128, 64, 128, "rand"), (256, 128, 256, "rand"), (300, 200, 512, "rand")]:
    rng = np.random.default_rng(M + N + K)
    if kind == "ints":
        AT = rng.integers(-8, 9, (K, M)).astype(np.float64)
        B = rng.integers(-8, 9, (K, N)).astype(np.float64)
    else:
        AT = rng.uniform(-1, 1, (K, M))
        B = rng.standard_normal((K, N)) * 6
    C = rng.standard_normal((M, N))
    r = cb.dbg.ozaki_gemm(AT, B, C, want_planes=True)
    ref, (pa, pb, ea, eb) = ozaki_ref.gemm(AT, B, C)
    print(f"case {M}x{N}x{K} {kind}: ea ok {np.array_equal(r['ea'], ea)} eb ok {np.array_equal(r['eb'], eb)} "
          f"planesA ok {np.array_equal(r['pa'], pa)} planesB ok {np.array_equal(r['pb'], pb)} "
          f"D bit-exact {np.array_equal(r['D'], ref)} max|D-ref| {np.abs(r['D'] - ref).max():.3e} "
          f"max|D-plain| {np.abs(r['D'] - (C - AT.T @ B)).max():.3e}  kernel {r['ms']:.3f} ms")
    if not np.array_equal(r["D"], ref):
        bad = np.argwhere(r["D"] != ref)
        print("   first mismatches (row, col):", bad[:8].tolist(), " count", len(bad), "of", M * N)
        print("   mismatching rows:", sorted(set(bad[:, 0].tolist()))[:16], " cols:", sorted(set(bad[:, 1].tolist()))[:16])
        print("   D[0,:4]", r["D"][0, :4], "ref", ref[0, :4], "C", C[0, :4])
