#!/usr/bin/env bash
# tools/gpu_session8.sh -- 8-GPU validation: 2x2x2 parity tests + the N=8 bench (BASELINE config C4)
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29519"
echo "== 2x2x2 tests"; timeout 150 python -u -m pytest tests/test_gpu_lu.py -m gpu -q --timeout 60 --timeout-method=thread -x -k "multi_gpu and (2-2-2 or golden)" > $O/s8_tests.log 2>&1; echo "rc=$?"; tail -2 $O/s8_tests.log
echo "== bench 8"; timeout 240 $T --nproc-per-node 8 bench.py --gpus 8 --steps 2 --warmup 3 > $O/s8_bench.log 2>$O/s8_bench.err; tail -c 1500 $O/s8_bench.log; tail -3 $O/s8_bench.err
