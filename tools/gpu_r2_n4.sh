#!/usr/bin/env bash
# round-2 4-GPU session: multi-rank parity (tests), bench variants, non-serialising timeline
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_cholesky.py tests/test_gpu_miniapp.py -x -q > gpurun_out/n4_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/n4_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 $TR bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/n4_bench_$name.log 2> gpurun_out/n4_bench_$name.err; echo "bench $name rc=$?"
}
run default CFLX_X=0
run dmma CFLX_GEMM=dmma
run cap32 CFLX_PANEL_CTAS=32
run dmma_cap32 CFLX_GEMM=dmma CFLX_PANEL_CTAS=32
timeout 300 $TR tools/timeline.py --gpus 4 --out gpurun_out/n4_timeline.json > gpurun_out/n4_timeline.log 2>&1; echo "timeline rc=$?"
CFLX_GEMM=dmma timeout 300 $TR tools/timeline.py --gpus 4 --out gpurun_out/n4_timeline_dmma.json >> gpurun_out/n4_timeline.log 2>&1; echo "timeline(ozaki) rc=$?"
timeout 600 $TR bench.py --algo cholesky --gpus 4 --steps 2 --warmup 3 > gpurun_out/n4_chol.log 2> gpurun_out/n4_chol.err; echo "chol bench rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/n4_bench_*.log")) + ["gpurun_out/n4_chol.log"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["parity"])
    except Exception as e:
        print(f, "failed", e)
PY
