#!/usr/bin/env bash
# round-2 session i: column-owner pivot search for <= 1024-row panels (isolated), then the final 1-GPU evidence
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "panel or column_owner" > gpurun_out/i_stack_tests.log 2>&1
rc=$?; echo "stack tests rc=$rc"; tail -5 gpurun_out/i_stack_tests.log
timeout 300 python tools/stack_speed.py > gpurun_out/i_stack_speed.log 2>&1; echo "stack speed rc=$?"; cat gpurun_out/i_stack_speed.log
if [ $rc -ne 0 ]; then echo "FALLING BACK to the row-owner kernel for the rest of the session"; export CFLX_STACK_KERNEL=0; fi
bash tools/gpu_r2_final1.sh
