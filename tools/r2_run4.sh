#!/bin/bash
# round 2, GPU call 4: Q3 persistent kernel with the fused M_x M_z plane stage (256-bit stores): full GPU suite, A/B timing
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest.log
{
for i in 1 2; do
timeout 120 python tools/prof_apply.py --n 64 --p 3
HPDG_B200_LIB=$PWD/dune-hpdg_b200/lib/libhpdg_b200_nofuse.so timeout 120 python tools/prof_apply.py --n 64 --p 3
done
timeout 120 python tools/prof_apply.py --n 96 --p 3
} > $O/r2d_timings.log 2>&1
tail -3 $O/r2d_pytest.log; cat $O/r2d_timings.log
