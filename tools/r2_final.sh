#!/bin/bash
# round 2, final single-GPU records: full GPU suite, bench lines (cfg2, cfg5), the other configs, launch list and ncu capture
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2z_pytest.log
timeout 900 python bench.py > $O/r2z_bench.json 2> $O/r2z_bench.err
timeout 900 python bench.py --workload cfg5 --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline > $O/r2z_bench_cfg5.json 2> $O/r2z_bench_cfg5.err
timeout 900 python tools/bench_configs.py --which cfg1,cfg3,cfg4small,cfg4,blockgs,nc > $O/r2z_configs.jsonl 2> $O/r2z_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2z_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q3_persist -s 2 -c 1 -o $O/r2z_q3p -f python tools/prof_apply.py --reps 2 > $O/r2z_ncu.log 2>&1
tail -3 $O/r2z_pytest.log; cut -c1-400 $O/r2z_bench.json; echo; cut -c1-300 $O/r2z_bench_cfg5.json; echo; cat $O/r2z_configs.jsonl; tail -2 $O/r2z_configs.err
