python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
export ISWM_BENCH_GRAPH=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/r2f_plain.log 2>&1 && tail -c 300 gpurun_out/r2f_plain.log &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 1700 -c 600 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu.log 2>&1
tail -n 2 gpurun_out/r2f_ncu.log
unset ISWM_BENCH_GRAPH
ncu --set full --clock-control none --import-source on -k regex:"tail_fused|tail_bwd|scale_crop_image|ccl_union|ccl_stats" -c 12 -o gpurun_out/r2f_prof_rows python tools/prof_tail.py > gpurun_out/r2f_ncu2.log 2>&1
tail -n 3 gpurun_out/r2f_ncu2.log; ls -la gpurun_out/ | tail -5
