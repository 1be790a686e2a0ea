#!/usr/bin/env bash
# tools/gpu_session4.sh -- 4-GPU validation: multi-rank parity tests (with and without look-ahead) + the N=4 bench
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
echo "== multi-gpu tests, look-ahead ON"; CFLX_LOOKAHEAD_MULTI=1 timeout 200 python -u -m pytest tests/test_gpu_lu.py -m gpu -q --timeout 80 --timeout-method=thread -x -k multi_gpu > $O/s4_tests_la.log 2>&1; echo "rc=$?"; tail -2 $O/s4_tests_la.log
echo "== bench 4, look-ahead OFF"; timeout 200 $T --nproc-per-node 4 bench.py --gpus 4 --steps 2 --warmup 3 > $O/s4_bench_off.log 2>$O/s4_bench_off.err; tail -c 1200 $O/s4_bench_off.log; tail -2 $O/s4_bench_off.err
echo "== bench 4, look-ahead ON"; CFLX_LOOKAHEAD_MULTI=1 timeout 200 $T --nproc-per-node 4 bench.py --gpus 4 --steps 2 --warmup 3 > $O/s4_bench_on.log 2>$O/s4_bench_on.err; tail -c 1200 $O/s4_bench_on.log; tail -2 $O/s4_bench_on.err
echo "== bench 2, look-ahead ON"; CFLX_LOOKAHEAD_MULTI=1 timeout 200 $T --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 3 > $O/s4_bench2_on.log 2>$O/s4_bench2_on.err; tail -c 1200 $O/s4_bench2_on.log; tail -2 $O/s4_bench2_on.err
