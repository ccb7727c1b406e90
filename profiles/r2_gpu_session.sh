#!/bin/bash
# One GPU session of round 2 (run through gpurun): (1) a quick parity subset with every round-1 kernel selected, then with one
# round-2 kernel switched on at a time; (2) the whole GPU suite with the defaults; (3) the micro-benchmarks; (4) bench.py.
mkdir -p gpurun_out
SUB='golden or single_vs_oracle or cluster_vs_oracle or align_batch or synthetic_single or k7_large_family_multicontig'
run_subset() {   # name, env...
  name=$1; shift
  env "$@" python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider -k "$SUB" > gpurun_out/r2_subset_$name.log 2>&1
  echo "subset $name: $(tail -1 gpurun_out/r2_subset_$name.log)"
}
run_subset old KGMA_PREFILTER=8mer KGMA_EVAL_KERNEL=serial KGMA_ALIGN_KERNEL=summary
run_subset tagged_only KGMA_PREFILTER=8mer KGMA_EVAL_KERNEL=serial
run_subset eval_only KGMA_PREFILTER=8mer KGMA_ALIGN_KERNEL=summary
run_subset prefilter9_only KGMA_EVAL_KERNEL=serial KGMA_ALIGN_KERNEL=summary
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest.log 2>&1
echo "full suite: $(tail -3 gpurun_out/r2_pytest.log | tr '\n' ' ')"
grep -E "^(FAILED|ERROR)" gpurun_out/r2_pytest.log | head -40
python profiles/r2_micro.py > gpurun_out/r2_micro.json 2> gpurun_out/r2_micro.err; tail -c 400 gpurun_out/r2_micro.err; cat gpurun_out/r2_micro.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 800 gpurun_out/r2_bench_n1.err; head -c 5000 gpurun_out/r2_bench_n1.json
