#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "tagged or align or golden or synthetic_single or cluster_vs_oracle or shard" > gpurun_out/r2d_pytest.log 2>&1
echo "subset: $(tail -3 gpurun_out/r2d_pytest.log | tr '\n' ' ')"
grep -E "^(FAILED|ERROR)|^E " gpurun_out/r2d_pytest.log | head -30
python profiles/r2_micro2.py > gpurun_out/r2d_micro2.json 2> gpurun_out/r2d_micro2.err; tail -c 600 gpurun_out/r2d_micro2.err; cat gpurun_out/r2d_micro2.json
python profiles/r2_micro3.py > gpurun_out/r2e_micro3.json 2> gpurun_out/r2e_micro3.err; tail -c 500 gpurun_out/r2e_micro3.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_micro3.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {n:(round(x['align_ms'],3),x['two_sweeps'],x['summary'],x['head']) for n,x in v.items()})
PY
