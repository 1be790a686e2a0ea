#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ozaki.py -x -q > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/g_tests.log
timeout 200 python tools/ozaki_speed.py > gpurun_out/g_speed.log 2>&1; echo "speed rc=$?"; cat gpurun_out/g_speed.log
CFLX_OZAKI_DBG=1 timeout 200 python tools/ozaki_speed.py 16128 16128 256 1 > gpurun_out/g_speed_dbg.log 2>&1; grep cycles gpurun_out/g_speed_dbg.log
timeout 120 python - > gpurun_out/g_panel.log 2>&1 <<'PY'
import sys, os; sys.path.insert(0, '.')
import numpy as np, conflux_b200 as cb
rng = np.random.default_rng(0)
for (n, v) in [(1024, 512), (2048, 512), (512, 256), (1024, 256), (16384, 256), (16384, 512)]:
    P = 5.0 + rng.random((n, v))
    _, _, _, ms = cb.dbg.panel(P, reps=3)
    print(f"panel n={n} v={v}: {ms:.3f} ms = {ms * 1e3 / v:.2f} us/col")
PY
cat gpurun_out/g_panel.log
CFLX_CLUSTER_ROWS=0 timeout 120 python - >> gpurun_out/g_panel.log 2>&1 <<'PY'
import sys, os; sys.path.insert(0, '.')
import numpy as np, conflux_b200 as cb
rng = np.random.default_rng(0)
for (n, v) in [(1024, 512), (2048, 512), (1024, 256)]:
    P = 5.0 + rng.random((n, v))
    _, _, _, ms = cb.dbg.panel(P, reps=3)
    print(f"[no cluster] panel n={n} v={v}: {ms:.3f} ms = {ms * 1e3 / v:.2f} us/col")
PY
tail -3 gpurun_out/g_panel.log
timeout 600 python -m pytest tests/test_gpu_lu.py -x -q > gpurun_out/g_lu.log 2>&1; echo "lu rc=$?"; tail -2 gpurun_out/g_lu.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_$name.log 2> gpurun_out/g_bench_$name.err; echo "bench $name rc=$?"; }
b default CFLX_X=1
b nocluster CFLX_CLUSTER_ROWS=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/g_bench_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["roofline"].get("achieved"), d["roofline"].get("frac"), d["parity"]["pivots_equal_reference"], d["parity"]["residual_PA_minus_LU_rel_frobenius"])
    except Exception as e:
        print(f, "failed", e)
PY
