#!/usr/bin/env bash
# round-2 final 1-GPU evidence session
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/z_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/z_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/z_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/z_bench_N1.log 2> gpurun_out/z_bench_N1.err; echo "bench rc=$?"
CFLX_GEMM=dmma timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/z_bench_N1_dmma.log 2> gpurun_out/z_bench_N1_dmma.err; echo "bench(dmma) rc=$?"
timeout 300 python tools/timeline.py --gpus 1 --out gpurun_out/z_timeline_N1.json > gpurun_out/z_timeline.log 2>&1; echo "timeline rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm -c 1 -o gpurun_out/r02_prof_ozaki_final python tools/ozaki_speed.py 16128 16128 256 1 > gpurun_out/z_ncu_ozaki.log 2>&1; echo "ncu ozaki rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 1 -c 1 -o gpurun_out/r02_prof_dmma_final python tools/ozaki_speed.py 16128 16128 256 1 > gpurun_out/z_ncu_dmma.log 2>&1; echo "ncu dmma rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py > gpurun_out/z_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_chol.csv python tools/profile_chol.py 16384 512 > gpurun_out/z_launches_chol.log 2>&1; echo "chol launch list rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/z_bench_N1.log", "gpurun_out/z_bench_N1_dmma.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["roofline"]["achieved"], d["roofline"]["frac"], d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
