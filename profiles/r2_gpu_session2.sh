#!/bin/bash
# Round 2, second GPU session: whole GPU suite, then every bench config at N=1 (short), then the reference arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1
echo "full suite: $(tail -3 gpurun_out/r2b_pytest.log | tr '\n' ' ')"
grep -E "^(FAILED|ERROR)" gpurun_out/r2b_pytest.log | head -40
for cfg in cluster exact k7; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_$cfg.json 2> gpurun_out/r2b_bench_$cfg.err
  echo "== $cfg rc=$?"; tail -c 600 gpurun_out/r2b_bench_$cfg.err; head -c 2500 gpurun_out/r2b_bench_$cfg.json; echo
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_ref_single.json 2> gpurun_out/r2b_ref_single.err
echo "== ref rc=$?"; tail -c 300 gpurun_out/r2b_ref_single.err; head -c 1500 gpurun_out/r2b_ref_single.json
