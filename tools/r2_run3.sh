#!/bin/bash
# round 2, GPU call 3: hp apply with the face-trace pass: parity tests, cfg3 timing, launch list
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "hp or high_order or cfg1 or uniform_3d or transfer or galerkin or vcycle or tuple or jacobi or smoke" > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.log
timeout 600 python tools/bench_configs.py --which cfg3,cfg1 > $O/r2c_cfg3.jsonl 2> $O/r2c_cfg3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2c_hp_launches.csv python tools/hp_once.py > $O/r2c_ncu.log 2>&1
tail -3 $O/r2c_pytest.log; cat $O/r2c_cfg3.jsonl; tail -2 $O/r2c_cfg3.err; python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r2c_hp_launches.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
from collections import defaultdict
d=defaultdict(list)
for r in rows[1:]:
    d[r[ki][:70]].append(float(r[vi]))
for k,v in d.items(): print(k, len(v), 'launches, mean %.1f us' % (sum(v)/len(v)/1e3 if max(v)>1000 else sum(v)/len(v)))
PY
