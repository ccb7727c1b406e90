#!/bin/bash
# whole GPU suite + the default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1
echo "full suite: $(tail -3 gpurun_out/r2i_pytest.log | tr '\n' ' ')"
grep -E "^(FAILED|ERROR)" gpurun_out/r2i_pytest.log | head -40
python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2i_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i_bench_n1.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "floor", d["e2e"]["h2d_ceiling"]["floor_ms_per_step"])
print(d["device_ms_per_step"]); print(d["host_ms_per_step"]); print(d["roofline"]["frac"], d["clocks"])
for k,v in d["other_configs"].items(): print(k, {a:b for a,b in v.items() if a in ("ms_per_step","value","device_ms","e2e")})
print(d["cpu_baseline"])
PY
