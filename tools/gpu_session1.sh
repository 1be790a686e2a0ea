#!/usr/bin/env bash
# tools/gpu_session1.sh -- single-GPU validation + variant measurements (run under gpurun, logs into gpurun_out/)
O=gpurun_out
echo "== tests (defaults)"; timeout 300 python -m pytest tests -m gpu -q --timeout 120 > $O/s1_tests.log 2>&1; echo "rc=$?"; tail -2 $O/s1_tests.log
echo "== panel+lu tests with the cluster exchange"; CFLX_CLUSTER_ROWS=6144 timeout 300 python -m pytest tests -m gpu -q --timeout 120 -k "panel or single_gpu" > $O/s1_tests_cluster.log 2>&1; echo "rc=$?"; tail -2 $O/s1_tests_cluster.log
echo "== gemm+lu tests with the 64x128 tile"; CFLX_GEMM_TILE=64 timeout 300 python -m pytest tests -m gpu -q --timeout 120 -k "gemm or trsm or single_gpu" > $O/s1_tests_tile64.log 2>&1; echo "rc=$?"; tail -2 $O/s1_tests_tile64.log
echo "== probe default"; timeout 200 python tools/probe.py > $O/s1_probe_default.log 2>&1; cat $O/s1_probe_default.log
echo "== probe tile64"; CFLX_GEMM_TILE=64 timeout 200 python tools/probe.py > $O/s1_probe_tile64.log 2>&1; grep -E "gemm_tn|LU " $O/s1_probe_tile64.log
echo "== probe cluster"; CFLX_CLUSTER_ROWS=6144 timeout 200 python tools/probe.py > $O/s1_probe_cluster.log 2>&1; grep -E "panel|LU " $O/s1_probe_cluster.log
echo "== probe tile64+cluster"; CFLX_GEMM_TILE=64 CFLX_CLUSTER_ROWS=6144 timeout 200 python tools/probe.py > $O/s1_probe_both.log 2>&1; grep -E "LU " $O/s1_probe_both.log
