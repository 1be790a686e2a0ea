#!/usr/bin/env bash
# round-2 last session (2 GPUs, what was left of the budget): N=2 line with the final code, Cholesky N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 90 $TR bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/n2b_bench.log 2> gpurun_out/n2b_bench.err; echo "bench rc=$?"
timeout 50 $TR bench.py --algo cholesky --gpus 2 --steps 2 --warmup 2 > gpurun_out/n2b_chol.log 2> gpurun_out/n2b_chol.err; echo "chol rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/n2b_bench.log", "gpurun_out/n2b_chol.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["parity"])
    except Exception as e:
        print(f, "failed", e)
PY
