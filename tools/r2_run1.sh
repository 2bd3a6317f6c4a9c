#!/bin/bash
# round 2, GPU call 1: full GPU test suite, timings of the uniform kernels, bench line, ncu capture of the headline kernel
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
{
for p in 3 4; do
  timeout 120 python tools/prof_apply.py --n 64 --p $p --jacobi 1
  timeout 120 python tools/prof_apply.py --n 64 --p $p --jacobi 1 --variant 40
done
timeout 120 python tools/prof_apply.py --n 64 --p 2 --jacobi 1
timeout 120 python tools/prof_apply.py --n 64 --p 1 --jacobi 1
timeout 120 python tools/prof_apply.py --n 32 --p 5 --jacobi 1
timeout 200 python tools/prof_apply.py --n 128 --p 4 --jacobi 1
} > $O/r2a_timings.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2a_bench.json 2> $O/r2a_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q3_persist -s 2 -c 1 -o $O/r2a_q3p -f python tools/prof_apply.py --reps 2 > $O/r2a_ncu.log 2>&1
tail -3 $O/r2a_pytest.log; cat $O/r2a_timings.log; cat $O/r2a_bench.json | cut -c1-600
