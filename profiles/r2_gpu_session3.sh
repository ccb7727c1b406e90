#!/bin/bash
# Round 2, profiling session: launch list of a short bench run, then ncu --set full of the round-2 kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 1 --no-cpu --no-extra --regions 1"
$CMD > gpurun_out/r2c_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps3.csv $CMD > gpurun_out/r2c_ncu1.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/r2c_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:'kgma_prefilter9|kgma_eval|kgma_align_tagged|kgma_align_summary' -s 8 -c 6 -f -o gpurun_out/r2_prof_kernels $CMD > gpurun_out/r2c_ncu2.log 2>&1
echo "set full rc=$?"; tail -3 gpurun_out/r2c_ncu2.log
ls -la gpurun_out/r2_prof_kernels.ncu-rep
