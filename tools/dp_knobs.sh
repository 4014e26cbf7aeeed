#!/bin/bash
# data-parallel transport knobs on N GPUs of one box: tools/dp_knobs.sh N "ENV=.." "ENV=.." ...
N=$1; shift
port=29600
for e in "$@"; do
  port=$((port+1))
  env $e timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | grep '^{"metric"' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('[$e]', d['n_gpus'], 'GPUs', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms/step')"
done
