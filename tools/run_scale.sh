#!/bin/bash
# usage: tools/run_scale.sh <n_gpus> <tag> <workload:steps> ...   -> gpurun_out/<tag>_<workload>_n<N>.json (+ .err)
# One torchrun bench per workload, each under its own timeout.
N=$1; TAG=$2; shift 2
PORT=29600
for ws in "$@"; do
  W=${ws%%:*}; S=${ws##*:}
  PORT=$((PORT+1))
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --workload $W --steps $S --warmup 3 > gpurun_out/${TAG}_${W}_n${N}.json 2> gpurun_out/${TAG}_${W}_n${N}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --workload $W --gpus $N --steps $S --warmup 3 > gpurun_out/${TAG}_${W}_n${N}.json 2> gpurun_out/${TAG}_${W}_n${N}.err
  fi
  echo "$W rc=$?" >> gpurun_out/${TAG}_status.log
done
