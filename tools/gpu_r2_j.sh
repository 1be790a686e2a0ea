#!/usr/bin/env bash
# round-2 session j: column-owner pivot search (refined bit-exactness test), blocked tile Cholesky, streamed e2e, CPU arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "panel or column_owner" > gpurun_out/j_stack_tests.log 2>&1
rc=$?; echo "stack tests rc=$rc"; tail -4 gpurun_out/j_stack_tests.log
timeout 300 python tools/stack_speed.py > gpurun_out/j_stack_speed.log 2>&1; echo "stack speed rc=$?"; cat gpurun_out/j_stack_speed.log
if [ $rc -ne 0 ]; then echo "FALLING BACK to the row-owner kernel for the rest of the session"; export CFLX_STACK_KERNEL=0; fi
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/j_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/j_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/j_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/j_bench_N1.log 2> gpurun_out/j_bench_N1.err; echo "bench rc=$?"
timeout 600 python bench.py --algo cholesky --N 16384 --steps 3 --warmup 2 > gpurun_out/j_chol_16k.log 2> gpurun_out/j_chol_16k.err; echo "chol bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/j_launches_chol.csv python tools/profile_chol.py 16384 512 > gpurun_out/j_launches_chol.log 2>&1; echo "chol launch list rc=$?"
CFLX_REF_BUDGET_S=30 timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/j_ref_arm.log 2> gpurun_out/j_ref_arm.err; echo "ref arm rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/j_bench_N1.log", "gpurun_out/j_chol_16k.log", "gpurun_out/j_ref_arm.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), d.get("e2e"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
