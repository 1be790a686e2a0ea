"""tools/ctx_probe.py -- does initialising torch's CUDA runtime between plan creation and factorisation break us?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import conflux_b200 as cb
mode = sys.argv[1]
if mode == "torch_first":
    import torch; torch.zeros(1, device="cuda")
comm = cb.Comm(1, 0, None, 0)
gv = cb.lu_params(1024, 1024, 128, 1, 1, 1, comm)
if mode == "torch_between":
    import torch; x = torch.empty(1024, pin_memory=True); torch.zeros(1, device="cuda")
perm = np.zeros(gv.M, dtype=np.int32)
try:
    print(mode, "ok", cb.LU_rep(gv, None, perm), sorted(perm.tolist()) == list(range(gv.M)))
except Exception as e:
    print(mode, "FAILED", e)
