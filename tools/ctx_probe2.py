import os, sys, ctypes, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import conflux_b200 as cb
from conflux_b200 import _lib
mode = sys.argv[1]
L = _lib.lib()
comm = cb.Comm(1, 0, None, 0)
gv = cb.lu_params(2048, 2048, 128, 1, 1, 1, comm)
perm = np.zeros(gv.M, dtype=np.int32)
cb.LU_rep(gv, None, perm)
cb.LU_rep(gv, None, None, upload=False)
p = None
if "smi" in mode:
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.DEVNULL)
    time.sleep(0.5)
if "timing" in mode:
    L.cflx_lu_set_kernel_timing(gv._h, 1)
if "barrier" in mode:
    comm.barrier()
try:
    for i in range(3):
        ms = cb.LU_rep(gv, None, None, upload=False)
        a, b = ctypes.c_double(), ctypes.c_double()
        L.cflx_lu_trailing_stats(gv._h, ctypes.byref(a), ctypes.byref(b))
    print(mode, "ok", ms, a.value)
except Exception as e:
    print(mode, "FAILED", e)
if p: p.terminate()
