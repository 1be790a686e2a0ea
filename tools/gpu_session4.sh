#!/usr/bin/env bash
# tools/gpu_session4.sh -- 4-GPU validation: multi-rank parity tests + the N=4 and N=2 bench lines
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
echo "== multi-gpu tests"; timeout 200 python -u -m pytest tests/test_gpu_lu.py -m gpu -q --timeout 80 --timeout-method=thread -x -k multi_gpu > $O/s4_tests.log 2>&1; echo "rc=$?"; tail -2 $O/s4_tests.log
echo "== bench 4"; timeout 200 $T --nproc-per-node 4 bench.py --gpus 4 --steps 2 --warmup 3 > $O/s4_bench4.log 2>$O/s4_bench4.err; tail -c 400 $O/s4_bench4.log; tail -2 $O/s4_bench4.err
echo "== bench 2"; timeout 200 $T --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 > $O/s4_bench2.log 2>$O/s4_bench2.err; tail -c 400 $O/s4_bench2.log; tail -2 $O/s4_bench2.err
