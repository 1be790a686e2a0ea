#!/usr/bin/env bash
mkdir -p gpurun_out
CFLX_OZAKI_DBG=1 timeout 200 python tools/ozaki_speed.py > gpurun_out/d_speed.log 2>&1; echo "speed rc=$?"; cat gpurun_out/d_speed.log
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_kernels.py -x -q > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/d_tests.log
timeout 120 python - > gpurun_out/d_panel.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, conflux_b200 as cb
rng = np.random.default_rng(0)
for (n, v) in [(16384, 256), (8192, 256), (1024, 512), (32768, 512), (16384, 512)]:
    P = 5.0 + rng.random((n, v))
    _, _, _, ms = cb.dbg.panel(P, reps=3)
    print(f"panel n={n} v={v}: {ms:.3f} ms = {ms * 1e3 / v:.2f} us/col")
PY
cat gpurun_out/d_panel.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -c 1 -o gpurun_out/r02_prof_ozaki python tools/ozaki_speed.py 16128 16128 256 1 > gpurun_out/d_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/d_ncu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench.log 2> gpurun_out/d_bench.err; echo "bench rc=$?"
CFLX_GEMM=ozaki timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_ozaki.log 2> gpurun_out/d_bench_ozaki.err; echo "bench(ozaki) rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/d_bench.log", "gpurun_out/d_bench_ozaki.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["pivots_equal_reference"], d["parity"]["residual_PA_minus_LU_rel_frobenius"])
    except Exception as e:
        print(f, "failed", e)
PY
