#!/usr/bin/env bash
# round-2 GPU session A: ozaki kernel bring-up + whole LU on the int8 path + timeline (1 GPU)
mkdir -p gpurun_out
CFLX_OZAKI_DBG=1 timeout 300 python -m pytest tests/test_gpu_ozaki.py -x -q -s > gpurun_out/a_ozaki.log 2>&1; echo "ozaki tests rc=$?"
CFLX_GEMM=ozaki timeout 600 python -m pytest tests/test_gpu_lu.py -x -q > gpurun_out/a_lu_ozaki.log 2>&1; echo "lu(ozaki) rc=$?"
CFLX_GEMM=ozaki timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_ozaki.log 2> gpurun_out/a_bench_ozaki.err; echo "bench(ozaki) rc=$?"
timeout 300 python tools/timeline.py --gpus 1 --out gpurun_out/a_timeline_N1.json > gpurun_out/a_timeline.log 2>&1; echo "timeline rc=$?"
CFLX_GEMM=ozaki timeout 300 python tools/timeline.py --gpus 1 --out gpurun_out/a_timeline_N1_ozaki.json >> gpurun_out/a_timeline.log 2>&1; echo "timeline(ozaki) rc=$?"
tail -4 gpurun_out/a_ozaki.log
