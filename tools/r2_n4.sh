#!/bin/bash
# N GPUs (N = 4: full set; N = 8: Krylov loop only): dist_check with the peer-memory halo, cfg2 bench line (driver's 20 steps), PCG loop
N=$1; WHAT=${2:-all}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$WHAT" = all ]; then
  HPDG_HALO=p2p timeout 900 $TR --master-port 29571 tools/dist_check.py > $O/r2_dist_check_n${N}_p2p.log 2>&1; echo "rc=$?" >> $O/r2_dist_check_n${N}_p2p.log
  grep "DIST_CHECK\|FAIL\|rc=" $O/r2_dist_check_n${N}_p2p.log | tail -3
  timeout 600 $TR --master-port 29572 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2p_bench_n${N}_s20.json 2> $O/r2p_bench_n${N}.err
  cut -c1-200 $O/r2p_bench_n${N}_s20.json
fi
timeout 600 $TR --master-port 29574 tools/pcg_bench.py > $O/r2q_pcg_n${N}.json 2> $O/r2q_pcg_n${N}.err
cat $O/r2q_pcg_n${N}.json; tail -2 $O/r2q_pcg_n${N}.err
