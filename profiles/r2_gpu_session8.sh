#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider -k "strobemer or golden or dists" > gpurun_out/r2l_pytest.log 2>&1
echo "subset: $(tail -3 gpurun_out/r2l_pytest.log | tr '\n' ' ')"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2l_pytest.log | head -30
KGMA_TRACE=1 python bench.py --config cluster --steps 2 --warmup 1 --no-cpu --no-extra --regions 1 > /dev/null 2> gpurun_out/r2l_trace_cluster.log; grep "kgma replay\|kgma scan\|kgma align" gpurun_out/r2l_trace_cluster.log | tail -30
