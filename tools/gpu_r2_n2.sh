#!/usr/bin/env bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/n2_bench.log 2> gpurun_out/n2_bench.err; echo "bench LU rc=$?"
timeout 600 $TR bench.py --algo cholesky --gpus 2 --steps 2 --warmup 3 > gpurun_out/n2_chol.log 2> gpurun_out/n2_chol.err; echo "bench chol rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/n2_bench.log", "gpurun_out/n2_chol.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["parity"])
    except Exception as e:
        print(f, "failed", e)
PY
