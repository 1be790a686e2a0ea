#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 120 python - > gpurun_out/c_umma.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import conflux_b200 as cb
for which, name in ((0, "kind::i8"), (1, "kind::f16 bf16")):
    for n in (64, 128, 256):
        t = cb.dbg.umma_peak(which, n)
        print(f"UMMA {name} 128x{n}: {t:.1f} TMAC/s = {2*t/1000:.3f} P(FL)OP/s")
PY
echo "umma rc=$?"; cat gpurun_out/c_umma.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/c_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/c_kernels.log
timeout 900 python -m pytest tests/test_gpu_lu.py -x -q > gpurun_out/c_lu.log 2>&1; echo "lu rc=$?"; tail -3 gpurun_out/c_lu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.log 2> gpurun_out/c_bench.err; echo "bench rc=$?"
CFLX_GEMM=ozaki timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_ozaki.log 2> gpurun_out/c_bench_ozaki.err; echo "bench(ozaki) rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c_bench.log", "gpurun_out/c_bench_ozaki.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["pivots_equal_reference"], d["parity"]["residual_PA_minus_LU_rel_frobenius"])
    except Exception as e:
        print(f, "failed", e)
PY
