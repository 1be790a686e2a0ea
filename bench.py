#!/usr/bin/env python
"""bench.py -- LU GFLOP/s (FP64, (2/3)N^3) of the CONFLUX hot path on 1/2/4/8 B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (N > 1: one rank per GPU)

A "step" is one LU factorisation of one synthetic matrix (lu_params::InitMatrix generator, seed 42):
    N=1: 16384 x 16384, v=256, grid 1x1x1   (BASELINE.json configs[1])
    N=2: 32768 x 32768, v=512, grid 1x1x2   (auto grid of 2 ranks, lu_params.hpp:21-47)
    N=4: 32768 x 32768, v=512, grid 2x2x1   (configs[2])
    N=8: 65536 x 65536, v=512, grid 2x2x2   (configs[3])
`value` times K steps with the matrix already resident in HBM (pristine device copy -> working copy -> factor),
`e2e` times K steps through the public LU_rep call with the host->device copy of the matrix from pinned memory
and the device->host read of the permutation inside the timed region.  PyTorch is plumbing only
(torch.distributed bootstrap, pinned host memory).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {1: (16384, 256, (1, 1, 1)), 2: (32768, 512, (1, 1, 2)), 4: (32768, 512, (2, 2, 1)),
             8: (65536, 512, (2, 2, 2))}
FP64_TENSOR_NOMINAL_TFLOPS = 40.0  # NVIDIA B200 datasheet (FP64 / FP64 tensor), context only


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        # median over the samples taken under load (upper half of the sorted clocks)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


CPU_THREAD_CAP = 32  # OpenBLAS (pthreads) + the reference's OpenMP loops oversubscribe badly beyond this


def cpu_threads():
    return max(1, min(os.cpu_count() or 1, CPU_THREAD_CAP))


_CPU_SNIPPET = r"""
import sys, time, json
sys.path.insert(0, %(root)r)
n, v, threads, reps = %(n)d, %(v)d, %(threads)d, %(reps)d
from oracle import ref
if ref.available():
    best = None
    for _ in range(reps):
        r = ref.lu_run(n, v, 1, 1, 1, n_rep=1, blas_threads=threads, want_factors=False)
        best = r["ms"] if best is None else min(best, r["ms"])
    print(json.dumps({"kind": "reference", "ms": best}))
else:
    from oracle import restate
    A = restate.init_matrix(n, v)
    t0 = time.time(); restate.lu(A, n, v)
    print(json.dumps({"kind": "port", "ms": (time.time() - t0) * 1e3}))
"""


def cpu_reference_run(n, v, threads, reps, timeout_s=240):
    """The reference's own CPU LU_rep (oracle/_ref = /root/reference sources + OpenBLAS; else the plain-C port) on a
    bounded sample, in a child process with a hard timeout so that a slow host cannot stall the GPU numbers."""
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS=str(threads))
    code = _CPU_SNIPPET % dict(root=ROOT, n=n, v=v, threads=threads, reps=reps)
    try:
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=timeout_s)
        d = json.loads(out.stdout.strip().splitlines()[-1])
        return d["kind"], d["ms"]
    except Exception as e:  # noqa: BLE001
        return "unavailable: %s" % type(e).__name__, None


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = cpu_threads()
    N, v, grid = WORKLOADS[args.gpus]
    n_s, v_s = 4096, 256
    t0 = time.time()
    cpu_reference_run(1024, 128, cores, max(args.warmup, 1))
    # one child process runs the K timed factorisations back to back (best-of is NOT taken: mean over K)
    tot_ms, kind, done = 0.0, "reference", 0
    for _ in range(args.steps):
        kind, ms = cpu_reference_run(n_s, v_s, cores, 1)
        if ms is None:
            break
        tot_ms += ms
        done += 1
    if done == 0:
        print(json.dumps({"impl": "reference", "unavailable": "CPU reference run failed or timed out (%s)" % kind}))
        return
    ms_step = tot_ms / done
    val = (2.0 / 3.0) * n_s ** 3 / (ms_step * 1e-3) / 1e9
    sample = f"N={n_s} v={v_s} grid 1x1x1, {args.steps} factorisation(s), OpenBLAS 0.3.15 ({cores} threads), no MPI/MKL in image"
    print(json.dumps({
        "impl": "reference", "metric": "LU GFLOP/s (FP64, (2/3)N^3)", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"LU N={N} v={v} grid {grid[0]}x{grid[1]}x{grid[2]}", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--N", type=int, default=0, help="override the matrix size (testing only; the line says so)")
    ap.add_argument("--v", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.gpus not in WORKLOADS or world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs WORLD_SIZE == {args.gpus} in {sorted(WORKLOADS)} (launch N>1 with torchrun)")
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    import conflux_b200 as cb
    from conflux_b200 import _lib

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    N, v, (Px, Py, Pz) = WORKLOADS[args.gpus]
    override = bool(args.N)
    if args.N:
        N = args.N
    if args.v:
        v = args.v
    L = _lib.lib()
    comm = cb.Comm.from_torch_distributed(device=local_rank) if world > 1 else cb.Comm(1, 0, None, local_rank)
    gv = cb.lu_params(N, N, v, Px, Py, Pz, comm)
    # pinned host staging for the end-to-end arm
    host = cb.pinned_empty((gv.Ml, gv.Nl))        # cudaHostAlloc through the library (torch never touches CUDA here)
    host[...] = gv.data
    gv.data = host
    perm = np.zeros(gv.M, dtype=np.int32)
    flops = (2.0 / 3.0) * float(gv.N) ** 3

    def barrier():
        comm.barrier()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # FP64 tensor-pipe peak of THIS GPU, measured before the runs (cold: burst) -- MEASURED_PEAKS.json has no FP64 entry
    peak_burst, peak_sustained = cb.dbg.fp64_peak_ex(0) if rank == 0 else (None, None)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident arm -------------------------------------------------------------------------------
    cb.LU_rep(gv, None, perm, upload=True)
    for _ in range(args.warmup - 1):
        cb.LU_rep(gv, None, None, upload=False)
    cnt = ctypes.c_int64()
    L.cflx_lu_launch_count(gv._h, ctypes.byref(cnt), 1)
    L.cflx_lu_set_kernel_timing(gv._h, 1)
    barrier()
    t0 = time.perf_counter()
    dev_ms, gemm_ms, gemm_flops = 0.0, 0.0, 0.0
    for _ in range(args.steps):
        dev_ms += cb.LU_rep(gv, None, None, upload=False)   # device-timed main loop (CUDA events on the rank's stream)
        a, b = ctypes.c_double(), ctypes.c_double()
        L.cflx_lu_trailing_stats(gv._h, ctypes.byref(a), ctypes.byref(b))
        gemm_ms += a.value
        gemm_flops += b.value
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    L.cflx_lu_launch_count(gv._h, ctypes.byref(cnt), 1)
    launches = cnt.value
    L.cflx_lu_set_kernel_timing(gv._h, 0)
    wall_ms = max_over_ranks(wall_ms)
    dev_ms = max_over_ranks(dev_ms)
    # the timed region is the reference's: the main loop between two grid barriers (conflux_opt.hpp:531-532,1805),
    # measured with CUDA events on each rank's launching stream, summed over the K steps, max over ranks
    ms_step = dev_ms / args.steps
    value = flops / (ms_step * 1e-3) / 1e9

    # ---- end-to-end arm: H2D of the matrix + factor + D2H of the permutation, every step -------------------------
    for _ in range(2):
        cb.LU_rep(gv, None, perm, upload=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cb.LU_rep(gv, None, perm, upload=True)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_val = flops / (e2e_ms * 1e-3) / 1e9

    # ---- parity of the run that was just timed ----------------------------------------------------------------
    ok_perm = sorted(perm.tolist()) == list(range(gv.M))
    resid = None
    if world == 1:
        try:
            resid = cb.residual(gv)                   # ||PA-LU||_F/||A||_F on the device, full BASELINE size
        except cb.ConfluxError as e:                  # noqa: F841
            resid = None

    if rank == 0:
        dmma_peak = peak_burst
        g_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        line = {
            "metric": "LU GFLOP/s (FP64, (2/3)N^3)", "value": value, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"LU N={gv.N} v={gv.v} grid {Px}x{Py}x{Pz}" + (" (size override, not the BASELINE config)" if override else ""),
                       "generator": "lu_params::InitMatrix mt19937_64(42+rank), 5+U[0,1)",
                       "l2": "inputs larger than L2 (local matrix %.1f GiB)" % (gv.Ml * gv.Nl * 8 / 2 ** 30),
                       "timing": "CUDA events on the launching stream around each factorisation's main loop, max over ranks",
                       "wall_ms_per_step_incl_restore_copy": wall_ms / args.steps},
            "e2e": {"value": e2e_val, "unit": "GFLOP/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(gv.Ml * gv.Nl * 8),
                    "d2h_bytes_per_step": int(gv.M * 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tn_kernel (trailing update, DMMA.8x8x4)",
                         "achieved": g_tf, "peak": dmma_peak, "unit": "TFLOP/s", "frac": (g_tf / dmma_peak) if g_tf else None,
                         # dram__bytes_read+write of ONE launch from the committed `ncu --set full` capture of the
                         # first-step shape (M=N=16128, K=256; profiles/r01_gemm_full.md): 2.798 + 2.058 GB, against
                         # 16*M*N + 8*K*(M+N) = 4.228 GB algorithmic (C read+write once, operands once)
                         "traffic": 4.856e9 if not override and args.gpus == 1 else None,
                         "traffic_launch": "M=N=16128 K=256 (step 0), algorithmic 4.228e9 B, ncu capture profiles/r01_gemm_full.md",
                         "peak_sustained": peak_sustained,
                         "peak_source": "FP64 tensor (DMMA.8x8x4) peak measured live on this GPU by cflx_dbg_fp64_peak_ex: `peak` = burst "
                                        "(best ~2 ms launch), `peak_sustained` = one 0.5 s launch under the power cap; "
                                        "MEASURED_PEAKS.json has no FP64 entry; nominal %.0f TFLOP/s" % FP64_TENSOR_NOMINAL_TFLOPS,
                         "share_of_step": gemm_ms / dev_ms if dev_ms else None,
                         "whole_path_frac": value / 1e3 / (args.gpus * dmma_peak)},
            "parity": {"permutation_is_permutation": bool(ok_perm), "residual_PA_minus_LU_rel_frobenius": resid,
                       "residual_tolerance": 1e-12},
        }
        if args.gpus == 1 and not args.no_cpu_baseline:
            cores = cpu_threads()
            n_s = 4096
            kind, ms = cpu_reference_run(n_s, 256, cores, 2)
            line["cpu_baseline"] = {"value": ((2.0 / 3.0) * n_s ** 3 / (ms * 1e-3) / 1e9) if ms else None, "unit": "GFLOP/s",
                                    "cores": cores, "kind": kind,
                                    "sample": f"LU_rep N={n_s} v=256 grid 1x1x1, best of 2, OpenBLAS 0.3.15 ({cores} threads of "
                                              f"{os.cpu_count()} cores); MKL/MPI are not in the image"}
        print(json.dumps(line))
    gv.free_comms()
    comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
