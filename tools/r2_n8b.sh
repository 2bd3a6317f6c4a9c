#!/bin/bash
# N GPUs: cfg2 bench with the driver's 20 steps and with 200 steps (p2p halo)
N=$1; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29572 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2p_bench_n${N}_s20.json 2> $O/r2p_bench_n${N}.err
timeout 600 $TR --master-port 29573 bench.py --gpus $N --steps 200 --warmup 10 > $O/r2p_bench_n${N}_s200.json 2>> $O/r2p_bench_n${N}.err
for f in $O/r2p_bench_n${N}_s20.json $O/r2p_bench_n${N}_s200.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value %.1f GDoF/s' % (d['value']/1e9), 'ms %.4f' % d['ms_per_step'], 'kernel_same %.4f' % d['roofline']['kernel_ms_same_buffers'], 'e2e %.2f' % (d['e2e']['value']/1e9), 'parity', d['parity']['rel_l2'])
except Exception as e:
    print(sys.argv[1], 'NO LINE', e)
PY
done
tail -3 $O/r2p_bench_n${N}.err
