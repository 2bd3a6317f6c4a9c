#!/bin/bash
# round 2, GPU call 2: Q3 persistent kernel with the interior-tile specialisation: parity subset, timing, instruction count
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "q3 or persistent or tuple or edge or vcycle or pcg or loop" > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
{
timeout 120 python tools/prof_apply.py --n 64 --p 3
timeout 120 python tools/prof_apply.py --n 64 --p 3
timeout 120 python tools/prof_apply.py --n 96 --p 3
} > $O/r2b_timings.log 2>&1
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:q3_persist -s 2 -c 1 python tools/prof_apply.py --reps 2 > $O/r2b_ncu.log 2>&1
tail -3 $O/r2b_pytest.log; cat $O/r2b_timings.log; grep -A8 "q3_persist" $O/r2b_ncu.log | tail -12
