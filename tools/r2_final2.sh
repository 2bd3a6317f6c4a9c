#!/bin/bash
# final check of the tree: smoke(), full GPU suite, default bench line, cfg4 record
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2y_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2y_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $O/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2y_pytest.log
timeout 900 python bench.py > $O/r2y_bench.json 2> $O/r2y_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2y_bench_ref.json 2> $O/r2y_bench_ref.err
timeout 900 python tools/bench_configs.py --which cfg3,cfg4small,cfg4 > $O/r2y_configs.jsonl 2> $O/r2y_configs.err
tail -2 $O/r2y_smoke.log; tail -3 $O/r2y_pytest.log; cut -c1-330 $O/r2y_bench.json; echo; cut -c1-300 $O/r2y_bench_ref.json; echo; cat $O/r2y_configs.jsonl | cut -c1-200
