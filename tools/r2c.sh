#!/bin/bash
tag=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_glue_gpu.py tests/test_unit_replay_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${tag}_tests.log; tail -5 gpurun_out/${tag}_tests.log
timeout 300 python tools/prof_shapes.py 3 > gpurun_out/${tag}_shapes.log 2>&1; cat gpurun_out/${tag}_shapes.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --profile-detail gpurun_out/${tag}_detail.txt > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-250
python - <<PY
import json
l=open('gpurun_out/${tag}_bench.log').read().strip().splitlines()[-1]
d=json.loads(l)
r=d['roofline']
print('igemm frac', r['frac'], 'ms', r['kernel_ms_per_step'])
for k,v in r['other_kernels'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
print('e2e', d['e2e']['value'])
PY
