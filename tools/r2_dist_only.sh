#!/bin/bash
# multi-GPU parity only: tools/r2_dist_only.sh N -> dist_check in both halo transports (uniform + distributed hp)
N=$1
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for halo in p2p nccl; do
  HPDG_HALO=$halo timeout 900 $TR --master-port 29571 tools/dist_check.py > $O/r2_dist_check_n${N}_${halo}.log 2>&1
  echo "rc=$?" >> $O/r2_dist_check_n${N}_${halo}.log
  grep "DIST_CHECK\|FAIL\|rc=\|hp p=" $O/r2_dist_check_n${N}_${halo}.log | tail -12
done
