#!/bin/bash
# same-box A/B of engine switches: tools/ab.sh "ENV=a ENV2=b" "ENV=c" ...   (each argument = one environment; "" = defaults)
for rep in 1 2; do
for e in "$@"; do
  env $e python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('[$e]', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms  igemm', round(r['kernel_ms_per_step'],3), round(r['frac'],4), {k:round(v['ms_per_step'],3) for k,v in r['other_kernels'].items()})"
done
done
