#!/bin/bash
# Round 2 closing session on the 8-GPU box: whole GPU suite, N=1 default line, ONE genome over 2 / 4 / 8 GPUs
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2v_pytest.log 2>&1; echo "suite: $(tail -1 gpurun_out/r2v_pytest.log)"
python bench.py > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err; echo "n=1 rc=$?"
for n in 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py --gpus $n > gpurun_out/r2v_bench_n$n.json 2> gpurun_out/r2v_bench_n$n.err; echo "n=$n rc=$?"; done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.load(open("gpurun_out/r2v_bench_n%d.json"%n))
        print(n, "res ms", round(d["ms_per_step"],3), round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"]), "floor", round(d["e2e"]["h2d_ceiling"]["floor_ms_per_step"],3), d.get("hits_equal_unsharded"), d["config"].get("shard"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e: print(n, "ERR", e)
PY
