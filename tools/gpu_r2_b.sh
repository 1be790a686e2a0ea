#!/usr/bin/env bash
mkdir -p gpurun_out
CFLX_OZAKI_DBG=1 timeout 300 python -m pytest tests/test_gpu_ozaki.py -x -q -s > gpurun_out/b_ozaki.log 2>&1; echo "ozaki tests rc=$?"
CFLX_GEMM=ozaki timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_ozaki.log 2> gpurun_out/b_bench_ozaki.err; echo "bench(ozaki) rc=$?"
grep -E "cycles|TFLOP" gpurun_out/b_ozaki.log | tail -4
