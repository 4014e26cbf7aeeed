#!/bin/bash
# One gpurun call: GPU tests, every bench workload, then the ncu passes (launch list + one --set full capture).
# usage: tools/gpu_round.sh <tag> [skip_ncu]
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python bench.py --steps 10 --warmup 3 --profile-detail gpurun_out/${tag}_detail.txt > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-400
python bench.py --workload lossmetric --steps 20 > gpurun_out/${tag}_lossmetric.log 2>&1; tail -1 gpurun_out/${tag}_lossmetric.log | cut -c1-1500
python bench.py --workload predict --steps 3 > gpurun_out/${tag}_predict.log 2>&1; tail -1 gpurun_out/${tag}_predict.log | cut -c1-600
python bench.py --backbone resnet101 --output-stride 8 --size 1024 --batch 4 --steps 3 --no-cpu-baseline > gpurun_out/${tag}_r101.log 2>&1; tail -1 gpurun_out/${tag}_r101.log | cut -c1-600
python tools/graph_probe.py > gpurun_out/${tag}_family_cost.log 2>&1; grep -v Warn gpurun_out/${tag}_family_cost.log | tail -4
python tools/kernel_floor.py > gpurun_out/${tag}_kernel_floor.log 2>&1
if [ -z "$2" ]; then
  export ISWM_BENCH_GRAPH=0      # ncu lists the eager launches (the graph replays the same kernels)
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/${tag}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 1700 -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 430 -c 3 -o gpurun_out/${tag}_prof_igemm $CMD > gpurun_out/${tag}_ncu2.log 2>&1
  tail -n 2 gpurun_out/${tag}_ncu.log; tail -n 2 gpurun_out/${tag}_ncu2.log
fi
