#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_cholesky.py -x -q > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/h_tests.log
timeout 200 python tools/ozaki_speed.py > gpurun_out/h_speed.log 2>&1; echo "speed rc=$?"; cat gpurun_out/h_speed.log
CFLX_OZAKI_DBG=1 timeout 200 python tools/ozaki_speed.py 16128 16128 256 1 > gpurun_out/h_speed_dbg.log 2>&1; grep cycles gpurun_out/h_speed_dbg.log
timeout 200 python tools/ozaki_speed.py 16384 16384 512 2 >> gpurun_out/h_speed.log 2>&1; tail -2 gpurun_out/h_speed.log
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_$name.log 2> gpurun_out/h_bench_$name.err; echo "bench $name rc=$?"; }
b default CFLX_X=1
timeout 600 python bench.py --algo cholesky --gpus 1 --steps 2 --warmup 3 --N 16384 --no-cpu-baseline > gpurun_out/h_chol_16k.log 2> gpurun_out/h_chol_16k.err; echo "chol bench rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/h_bench_*.log")) + ["gpurun_out/h_chol_16k.log"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["roofline"].get("achieved"), d["roofline"].get("frac"))
    except Exception as e:
        print(f, "failed", e)
PY
