#!/usr/bin/env bash
# round-2 session k (1 GPU, last one): full GPU test-suite, final N=1 bench line, SM split of the look-ahead pivot search at
# N=1, Cholesky with the shared-memory 128-block kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/k_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/k_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/k_bench_N1.log 2> gpurun_out/k_bench_N1.err; echo "bench rc=$?"
for cap in 48 64; do
  CFLX_PANEL_CTAS=$cap timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_cap$cap.log 2> gpurun_out/k_bench_cap$cap.err; echo "bench cap$cap rc=$?"
done
for leave in 8 24; do
  CFLX_CHOL_LEAVE=$leave timeout 200 python bench.py --algo cholesky --N 16384 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/k_chol_leave$leave.log 2> gpurun_out/k_chol_leave$leave.err; echo "chol leave$leave rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/k_launches_chol.csv python tools/profile_chol.py 16384 512 > gpurun_out/k_launches_chol.log 2>&1; echo "chol launch list rc=$?"
python - <<'PY'
import json, glob
for f in ["gpurun_out/k_bench_N1.log"] + sorted(glob.glob("gpurun_out/k_bench_cap*.log")) + sorted(glob.glob("gpurun_out/k_chol_leave*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
