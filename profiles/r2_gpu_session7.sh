#!/bin/bash
# Round 2: the other BASELINE configs sharded over N GPUs (run with gpurun --gpus 8), plus their N=1 lines
mkdir -p gpurun_out
run() { n=$1; cfg=$2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$3 bench.py --gpus $n --config $cfg --steps 20 --warmup 5 --no-extra > gpurun_out/r2k_bench_${cfg}_n$n.json 2> gpurun_out/r2k_bench_${cfg}_n$n.err; echo "$cfg n=$n rc=$?"; }
run 4 single 1; run 8 cluster 2; run 8 exact 3; run 8 k7 4
for c in cluster k7; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/r2k_bench_${c}_n1.json 2> gpurun_out/r2k_bench_${c}_n1.err; echo "$c n=1 rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2k_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "res ms", round(d["ms_per_step"],3), "value", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"]), "floor", round(d["e2e"]["h2d_ceiling"]["floor_ms_per_step"],3), d["device_ms_per_step"]["prefilter_or_match"], d["device_ms_per_step"]["count_table"], d["device_ms_per_step"]["extension"], d.get("hits_equal_unsharded"), d["hits_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
