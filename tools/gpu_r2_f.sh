#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python tools/ozaki_speed.py > gpurun_out/f_speed.log 2>&1; echo "speed rc=$?"; cat gpurun_out/f_speed.log
CFLX_OZAKI_DBG=1 timeout 200 python tools/ozaki_speed.py 16128 16128 256 1 > gpurun_out/f_speed_dbg.log 2>&1; grep cycles gpurun_out/f_speed_dbg.log
timeout 200 python tools/ozaki_speed.py 16384 16384 512 2 >> gpurun_out/f_speed.log 2>&1; tail -2 gpurun_out/f_speed.log
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_cholesky.py -x -q > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/f_tests.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_$name.log 2> gpurun_out/f_bench_$name.err; echo "bench $name rc=$?"; }
b dmma CFLX_X=1
b ozaki CFLX_GEMM=ozaki
timeout 600 python bench.py --algo cholesky --gpus 1 --steps 2 --warmup 3 --N 16384 --no-cpu-baseline > gpurun_out/f_chol_16k.log 2> gpurun_out/f_chol_16k.err; echo "chol bench rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/f_bench_*.log")) + ["gpurun_out/f_chol_16k.log"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["roofline"].get("achieved"), d["parity"])
    except Exception as e:
        print(f, "failed", e)
PY
