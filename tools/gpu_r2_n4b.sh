#!/usr/bin/env bash
# round-2 second 4-GPU session: column-owner tournament kernel A/B, SM split between look-ahead pivot search and trailing
# update, look-ahead off, blocked tile Cholesky (all at C3 = N=32768 v=512 2x2x1, same box)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 $TR bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/n4b_bench_$name.log 2> gpurun_out/n4b_bench_$name.err; echo "bench $name rc=$?"
}
run default CFLX_X=0
run rowowner CFLX_STACK_KERNEL=0
run cap64 CFLX_PANEL_CTAS=64
run cap96 CFLX_PANEL_CTAS=96
run nolookahead CFLX_LOOKAHEAD_MULTI=0
timeout 300 $TR tools/timeline.py --gpus 4 --out gpurun_out/n4b_timeline.json > gpurun_out/n4b_timeline.log 2>&1; echo "timeline rc=$?"
timeout 600 $TR bench.py --algo cholesky --gpus 4 --steps 2 --warmup 3 > gpurun_out/n4b_chol.log 2> gpurun_out/n4b_chol.err; echo "chol bench rc=$?"
timeout 600 python -m pytest tests/test_gpu_lu.py tests/test_gpu_cholesky.py -x -q -k "multi" > gpurun_out/n4b_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/n4b_tests.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/n4b_bench_*.log")) + ["gpurun_out/n4b_chol.log"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["parity"])
    except Exception as e:
        print(f, "failed", e)
PY
