"""Test helpers: run an SPMD function on P ranks as threads of this process (one GPU per rank), and the numpy
residual / assembly used by the parity tests."""
import threading

import numpy as np

import conflux_b200 as cb
from oracle import layout


def n_gpus():
    import ctypes
    n = ctypes.c_int()
    cb._lib.lib().cflx_device_count(ctypes.byref(n))
    return n.value


def run_ranks(P, fn):
    """fn(comm) on P threads; returns list of results in rank order; re-raises the first exception."""
    uid = cb.Comm.unique_id() if P > 1 else None
    out, err = [None] * P, [None] * P

    def body(r):
        try:
            comm = cb.Comm(P, r, uid, r)
            try:
                out[r] = fn(comm)
            finally:
                comm.close()
        except BaseException as e:  # noqa: BLE001
            err[r] = e
            import sys
            import traceback
            sys.stderr.write(f"[rank {r}] failed: {e!r}\n")
            traceback.print_exc()
            sys.stderr.flush()

    th = [threading.Thread(target=body, args=(r,)) for r in range(P)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


def gpu_lu(N, v, Px=1, Py=1, Pz=1, A_locals=None, want_C=True, want_resid=False):
    """Runs the CUDA path on a Px x Py x Pz grid.  Returns dict(A=[...], C=[...], perm, ms, dims)."""
    P = Px * Py * Pz

    def body(comm):
        gv = cb.lu_params(N, N, v, Px, Py, Pz, comm)
        if A_locals is not None:
            gv.data[...] = np.asarray(A_locals[gv.rank]).reshape(gv.Ml, gv.Nl)
        C = np.zeros((gv.Ml, gv.Nl)) if want_C else None
        perm = np.full(gv.M, -1, dtype=np.int32)
        ms = cb.LU_rep(gv, C, perm)
        res = dict(A=gv.data.copy(), C=C, perm=perm, ms=ms, rank=gv.rank)
        if want_resid:
            res["resid_abs"], res["resid"] = cb.validate(gv)              # collective, on the GPU grid
        gv.free_comms()
        return res

    rs = run_ranks(P, body)
    d = layout.dims(N, v, Px, Py, Pz)
    return dict(A=[r["A"] for r in rs], C=[r["C"] for r in rs], perm=rs[0]["perm"], perms=[r["perm"] for r in rs],
                ms=max(r["ms"] for r in rs), dims=d, resid=[r.get("resid") for r in rs],
                resid_abs=[r.get("resid_abs") for r in rs])
