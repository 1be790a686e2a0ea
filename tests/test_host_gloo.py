"""CPU, world_size 2 over gloo: the host-side plumbing of the N>1 path that does not need a GPU -- NCCL id bootstrap
through torch.distributed, rank -> grid coordinates, per-rank InitMatrix seeds/layers, identical auto-grid on all ranks,
and the loud refusal of the device entry point on a GPU-less box."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import conflux_b200 as cb
    out = {}
    # 1. id bootstrap: rank 0's id reaches every rank unchanged
    obj = [cb.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    out["id"] = obj[0]
    # 2. every rank derives the same grid and its own coordinates (rank = (pi*Py + pj)*Pz + pk)
    Px, Py, Pz = cb.auto_grid(4096, 4096, world)
    out["grid"] = (Px, Py, Pz)
    out["coords"] = (rank // (Py * Pz), (rank // Pz) % Py, rank % Pz)
    # 3. generator: layer 0 seeded with 42 + rank, other layers zero (lu_params.hpp:149-155,364-375)
    a = cb.init_matrix_host(64, 64, 16, Px, Py, Pz, rank)
    out["sum"] = float(a.sum())
    out["nz"] = bool(a.any())
    # 4. no CPU fallback: creating the communicator on a GPU-less box must fail loudly
    try:
        cb.Comm.from_torch_distributed(device=0)
        out["comm"] = "created"
    except cb.ConfluxError as e:
        out["comm"] = str(e)
    dist.barrier()
    q.put((rank, out))
    dist.destroy_process_group()


def test_two_rank_host_plumbing():
    import torch.multiprocessing as mp   # (torch is imported lazily: GPU-marked collection must stay fast)
    import conflux_b200 as cb
    import ctypes
    n = ctypes.c_int()
    cb._lib.lib().cflx_device_count(ctypes.byref(n))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["id"] == res[1]["id"] and len(res[0]["id"]) == 128
    assert res[0]["grid"] == res[1]["grid"] == (1, 1, 2)          # get_p_grid(P=2) (lu_params.hpp:21-47)
    assert res[0]["coords"] == (0, 0, 0) and res[1]["coords"] == (0, 0, 1)
    assert res[0]["nz"] and not res[1]["nz"]                       # layer 1 starts at zero
    from oracle import restate
    assert abs(res[0]["sum"] - restate.init_matrix(64, 16, 1, 1, 2)[0].sum()) < 1e-9
    if n.value == 0:
        for r in (0, 1):
            assert "no CPU fallback" in res[r]["comm"]
