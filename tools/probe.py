"""tools/probe.py -- quick single-GPU performance probes of the individual kernels (run under gpurun)."""
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import conflux_b200 as cb

print("FP64 pipe: DMMA %.2f TFLOP/s, DFMA %.2f TFLOP/s" % (cb.dbg.fp64_peak(0), cb.dbg.fp64_peak(1)), flush=True)
rng = np.random.default_rng(0)
for (M, N, K) in [(16128, 16128, 256), (8192, 8192, 512)]:
    AT, B = rng.standard_normal((K, M)), rng.standard_normal((K, N))
    _, ms = cb.dbg.gemm_tn(AT, B, None, -1.0, 1.0, reps=5)
    print(f"gemm_tn M={M} N={N} K={K}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.2f} TFLOP/s", flush=True)
for (n, v) in [(2048, 256), (4096, 256), (16384, 256), (32768, 512), (1024, 512)]:
    P = 5.0 + rng.random((n, v))
    _, _, _, ms = cb.dbg.panel(P, reps=3)
    import ctypes
    cyc = (ctypes.c_longlong * 8)()
    cb._lib.lib().cflx_dbg_last_panel_cycles(cyc)
    names = ["cand+argmax1", "exchange", "argmax2", "rowfetch", "eliminate", "load/wb", "u12", "update"]
    print(f"panel n={n} v={v}: {ms:.3f} ms ({ms/v*1e3:.2f} us/column)  CTA0 kcycles: " +
          ", ".join(f"{a}={c/1e3:.0f}" for a, c in zip(names, cyc)), flush=True)
comm = cb.Comm(1, 0, None, 0)
cfgs = [(1, 32, 8192, 256), (1, 32, 16384, 256)]
for (la, ctas, N, v) in cfgs:
    os.environ["CFLX_LOOKAHEAD"] = str(la)
    os.environ["CFLX_PANEL_CTAS"] = str(ctas)
    gv = cb.lu_params(N, N, v, 1, 1, 1, comm)
    perm = np.zeros(gv.M, dtype=np.int32)
    cb.LU_rep(gv, None, perm)
    ts = [cb.LU_rep(gv, None, None, upload=False) for _ in range(3)]
    ms = min(ts)
    cb._lib.lib().cflx_lu_set_profiling(gv._h, 1)
    cb.LU_rep(gv, None, None, upload=False)
    import ctypes
    ph = (ctypes.c_double * 8)()
    cb._lib.lib().cflx_lu_phase_ms(gv._h, ph)
    cb._lib.lib().cflx_lu_set_profiling(gv._h, 0)
    print(f"LU lookahead={la} panel_ctas={ctas} N={N} v={v}: {ms:.2f} ms  {2/3*N**3/ms/1e9:.2f} TFLOP/s   phases(ms, serialised): " +
          ", ".join(f"{n}={x:.1f}" for n, x in zip(["panel", "tourn", "moves", "reduce", "trsm", "gemm", "store", "other"], ph)),
          flush=True)
    gv.free_comms()
