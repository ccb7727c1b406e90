#!/bin/bash
# Round 2 final profiling session: launch list of a short bench run, ncu --set full of the kernels of two RESIDENT steps
# (the first 30 matching launches belong to the cold streamed call: chunk prefilters, count-table and extension launches),
# and a KGMA_TRACE=1 run for the host-side phase times.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 1 --no-cpu --no-extra --regions 1"
$CMD > gpurun_out/r2j_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps3.csv $CMD > gpurun_out/r2j_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'kgma_prefilter9|kgma_eval|kgma_align_tagged' -s 30 -c 6 -f -o gpurun_out/r2_prof_kernels $CMD > gpurun_out/r2j_ncu2.log 2>&1
echo "set full rc=$?"
KGMA_TRACE=1 python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --regions 1 > /dev/null 2> gpurun_out/r2j_trace.log; tail -40 gpurun_out/r2j_trace.log
