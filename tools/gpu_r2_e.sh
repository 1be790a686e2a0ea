#!/usr/bin/env bash
mkdir -p gpurun_out
CFLX_OZAKI_DBG=1 timeout 200 python tools/ozaki_speed.py > gpurun_out/e_speed.log 2>&1; echo "speed rc=$?"; cat gpurun_out/e_speed.log
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_kernels.py -x -q > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/e_tests.log
CFLX_PANEL_SPEC=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lu.py -x -q > gpurun_out/e_tests_spec.log 2>&1; echo "tests(spec) rc=$?"; tail -2 gpurun_out/e_tests_spec.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-validate > gpurun_out/e_bench_$name.log 2> gpurun_out/e_bench_$name.err; echo "bench $name rc=$?"; }
b dmma_nb64 CFLX_TRSM_NB=64
b dmma_nb128 CFLX_TRSM_NB=128
b dmma_nb128_spec CFLX_TRSM_NB=128 CFLX_PANEL_SPEC=1
b dmma_nb64_b CFLX_TRSM_NB=64
b ozaki_nb128 CFLX_GEMM=ozaki CFLX_TRSM_NB=128
b ozaki_nb128_spec CFLX_GEMM=ozaki CFLX_TRSM_NB=128 CFLX_PANEL_SPEC=1
b ozaki_nb64 CFLX_GEMM=ozaki CFLX_TRSM_NB=64
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/e_bench_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 2), round(d["roofline"]["achieved"], 2), d["parity"]["pivots_equal_reference"])
    except Exception as e:
        print(f, "failed", e)
PY
