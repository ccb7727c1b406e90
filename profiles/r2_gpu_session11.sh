#!/bin/bash
# Round 2, final profiling pass on the final tree: launch list, ncu --set full of two resident steps' kernels, and of the dense pass
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 1 --no-cpu --no-extra --regions 1"
$CMD > gpurun_out/r2y_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps3_final.csv $CMD > gpurun_out/r2y_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'kgma_prefilter9|kgma_eval|kgma_align_tagged' -s 100 -c 6 -f -o gpurun_out/r2_prof_kernels_final $CMD > gpurun_out/r2y_ncu2.log 2>&1
echo "set full rc=$?"
python profiles/r2_dense_once.py > gpurun_out/r2y_dense_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:kgma_eval -s 1 -c 1 -f -o gpurun_out/r2_prof_dense_final python profiles/r2_dense_once.py > gpurun_out/r2y_ncu3.log 2>&1
echo "dense rc=$?"; cat gpurun_out/r2y_dense_plain.log
