#!/bin/bash
# round-2 first call: baseline tests + bench, then ncu --set full of the time-losing shapes
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2a_tests.log 2>&1; tail -3 gpurun_out/r2a_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.log 2>&1; tail -1 gpurun_out/r2a_bench.log | cut -c1-300
python tools/prof_shapes.py 3 > gpurun_out/r2a_shapes.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_igemm|conv_wgrad|bn_bwd' -o gpurun_out/r2a_shapes python tools/prof_shapes.py 1 > gpurun_out/r2a_ncu.log 2>&1
cat gpurun_out/r2a_shapes.log; tail -3 gpurun_out/r2a_ncu.log
