#!/usr/bin/env bash
# round-2 8-GPU session: multi-rank parity (8-GPU test cases), LU C4 bench + timeline, Cholesky C5 bench + miniapp
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_cholesky.py -x -q -k "multi or golden" > gpurun_out/n8_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/n8_tests.log
timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/n8_bench.log 2> gpurun_out/n8_bench.err; echo "bench LU rc=$?"
timeout 600 $TR bench.py --algo cholesky --gpus 8 --steps 3 --warmup 3 > gpurun_out/n8_chol.log 2> gpurun_out/n8_chol.err; echo "bench chol rc=$?"
timeout 300 $TR tools/timeline.py --gpus 8 --out gpurun_out/n8_timeline.json > gpurun_out/n8_timeline.log 2>&1; echo "timeline rc=$?"
make -C examples > /dev/null 2>&1
timeout 300 ./examples/cholesky_miniapp --dim 32768 --tile 512 --run 2 --ranks 8 --validate > gpurun_out/n8_chol_miniapp.log 2>&1; echo "chol miniapp rc=$?"; tail -12 gpurun_out/n8_chol_miniapp.log
timeout 300 ./examples/conflux_miniapp -N 32768 -b 512 -r 1 -p 2,2,2 --validate > gpurun_out/n8_lu_miniapp.log 2>&1; echo "lu miniapp rc=$?"; tail -4 gpurun_out/n8_lu_miniapp.log
python - <<'PY'
import json
for f in ("gpurun_out/n8_bench.log", "gpurun_out/n8_chol.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["parity"])
    except Exception as e:
        print(f, "failed", e)
PY
