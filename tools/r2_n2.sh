#!/bin/bash
# N = 2 check of the halo-pack change: parity (p2p), bench cfg2 with 20 and 200 steps
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests -m gpu -q -x -k "two_rank or two_ranks" 2>&1 | tail -2
timeout 600 $TR --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2o_bench_n2_s20.json 2> $O/r2o_bench_n2.err
timeout 600 $TR --master-port 29573 bench.py --gpus 2 --steps 200 --warmup 10 > $O/r2o_bench_n2_s200.json 2>> $O/r2o_bench_n2.err
for f in $O/r2o_bench_n2_s20.json $O/r2o_bench_n2_s200.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value %.1f GDoF/s' % (d['value']/1e9), 'ms %.4f' % d['ms_per_step'], 'kernel_same %.4f' % d['roofline']['kernel_ms_same_buffers'], 'e2e %.2f' % (d['e2e']['value']/1e9), 'parity', d['parity']['rel_l2'])
except Exception as e:
    print(sys.argv[1], 'NO LINE', e)
PY
done
tail -3 $O/r2o_bench_n2.err
