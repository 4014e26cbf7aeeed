#!/bin/bash
# Round-2 evidence pass, ONE gpurun call: usage tools/r2_evidence.sh <tag>
#   1. every bench workload without a profiler (the only numbers that count), per-launch detail, family marginal costs
#   2. ncu launch list of the eager step (same kernels as the graph replay), 3. ncu --set full of the kernel shapes that lose the
#   most time (convolution / weight gradient / BatchNorm reduce) and of the cfg5 loss / metric kernels (DRAM traffic)
tag=$1
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --profile-detail gpurun_out/${tag}_detail.txt > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-300
python bench.py --workload lossmetric --steps 20 > gpurun_out/${tag}_lossmetric.log 2>&1; tail -1 gpurun_out/${tag}_lossmetric.log | cut -c1-600
python bench.py --workload predict --steps 3 > gpurun_out/${tag}_predict.log 2>&1; tail -1 gpurun_out/${tag}_predict.log | cut -c1-400
python bench.py --backbone resnet101 --output-stride 8 --size 1024 --batch 4 --steps 3 --no-cpu-baseline > gpurun_out/${tag}_r101.log 2>&1; tail -1 gpurun_out/${tag}_r101.log | cut -c1-300
python tools/graph_probe.py 2>&1 | grep -v Warn > gpurun_out/${tag}_family_cost.log; tail -6 gpurun_out/${tag}_family_cost.log
python tools/kernel_floor.py > gpurun_out/${tag}_kernel_floor.log 2>&1
python tools/prof_wgrad.py 3 > gpurun_out/${tag}_wgrad_shapes.log 2>&1
python tools/prof_shapes.py 2 > gpurun_out/${tag}_shapes.log 2>&1
export ISWM_BENCH_GRAPH=0      # ncu lists the eager launches (the graph replays the same kernels)
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 1600 -c 560 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -n 1 gpurun_out/${tag}_ncu.log
ncu --set full --clock-control none --import-source on -k regex:'conv_igemm|conv_wgrad|bn_bwd_reduce' -c 12 -o gpurun_out/${tag}_shapes python tools/prof_shapes.py 1 > gpurun_out/${tag}_ncu2.log 2>&1
tail -n 1 gpurun_out/${tag}_ncu2.log
unset ISWM_BENCH_GRAPH
python bench.py --workload lossmetric --steps 2 > /dev/null 2>&1 &&
for k in class_hist wce2 argmax_confusion; do
  ncu --set full --clock-control none -k regex:$k -c 1 -o gpurun_out/${tag}_lossmetric_$k python bench.py --workload lossmetric --steps 1 > gpurun_out/${tag}_ncu3_$k.log 2>&1
  tail -n 1 gpurun_out/${tag}_ncu3_$k.log
done
cuobjdump -sass iswm_b200/libiswm_b200.so > /tmp/sass.txt 2>/dev/null; python tools/sass_summary.py /tmp/sass.txt > gpurun_out/${tag}_sass.txt 2>&1; head -30 gpurun_out/${tag}_sass.txt
