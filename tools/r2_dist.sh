#!/bin/bash
# round 2 multi-GPU evidence: tools/r2_dist.sh N  -> dist_check (oracle parity on the GLOBAL mesh: apply, chained applies, dot,
# block Jacobi, V-cycle) in both halo transports, then the cfg2 and cfg5 bench lines at N GPUs
N=$1
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for halo in p2p nccl; do
  HPDG_HALO=$halo timeout 900 $TR --master-port 29571 tools/dist_check.py > $O/r2_dist_check_n${N}_${halo}.log 2>&1
  echo "rc=$?" >> $O/r2_dist_check_n${N}_${halo}.log
  grep "DIST_CHECK\|FAIL\|rc=" $O/r2_dist_check_n${N}_${halo}.log | tail -3
done
timeout 900 $TR --master-port 29572 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_cfg2_n${N}.json 2> $O/r2_bench_cfg2_n${N}.err
timeout 900 $TR --master-port 29573 bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 --e2e-steps 3 > $O/r2_bench_cfg5_n${N}.json 2> $O/r2_bench_cfg5_n${N}.err
HPDG_HALO=nccl timeout 900 $TR --master-port 29574 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_cfg2_n${N}_nccl.json 2> $O/r2_bench_cfg2_n${N}_nccl.err
for f in $O/r2_bench_cfg2_n${N}.json $O/r2_bench_cfg5_n${N}.json $O/r2_bench_cfg2_n${N}_nccl.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value %.1f GDoF/s' % (d['value']/1e9), 'ms %.4f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.2f' % (d['e2e']['value']/1e9), 'parity', d['parity']['rel_l2'])
except Exception as e:
    print(sys.argv[1], 'NO LINE', e)
PY
done
tail -q -n 3 $O/r2_bench_cfg2_n${N}.err $O/r2_bench_cfg5_n${N}.err; true
