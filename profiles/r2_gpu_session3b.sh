#!/bin/bash
# ncu --set full of the kernels of two RESIDENT steps (the first 31 matching launches belong to the cold streamed call)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 1 --no-cpu --no-extra --regions 1"
$CMD > gpurun_out/r2c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'kgma_prefilter9|kgma_eval|kgma_align_tagged|kgma_align_summary' -s 31 -c 8 -f -o gpurun_out/r2_prof_kernels $CMD > gpurun_out/r2c_ncu2.log 2>&1
echo "set full rc=$?"; tail -2 gpurun_out/r2c_ncu2.log | cut -c1-300
