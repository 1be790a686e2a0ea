#!/usr/bin/env bash
# tools/gpu_session_final.sh -- the evidence run: bench line + ncu captures that profiles/ summarises
O=gpurun_out
echo "== bench N=1"; timeout 200 python bench.py --steps 5 --warmup 3 > $O/f_bench.log 2>$O/f_bench.err; tail -c 300 $O/f_bench.log; tail -2 $O/f_bench.err
echo "== ncu full: trailing-update GEMM alone (16128^2 x 256)"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 1 -c 1 -o $O/f_prof_gemm python tools/profile_gemm.py > $O/f_ncu_gemm.log 2>&1; tail -1 $O/f_ncu_gemm.log
echo "== ncu full: panel kernel (step 3 of N=8192)"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:panel_getrf -s 3 -c 1 -o $O/f_prof_panel python tools/profile_step.py 8192 256 > $O/f_ncu_panel.log 2>&1; tail -1 $O/f_ncu_panel.log
echo "== ncu launch list of one factorisation (N=16384, v=256)"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/f_launches.csv python tools/profile_step.py > $O/f_ncu_list.log 2>&1; tail -1 $O/f_ncu_list.log
