#!/bin/bash
python tools/prof_shapes.py 2 > gpurun_out/r2i_shapes.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_igemm' -c 6 -o gpurun_out/r2i_shapes python tools/prof_shapes.py 1 > gpurun_out/r2i_ncu.log 2>&1
tail -2 gpurun_out/r2i_ncu.log
