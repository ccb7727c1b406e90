"""one-off extended fuzz: run tests/test_gpu_parity.py::test_fuzz_vs_oracle for many more seeds"""
import sys, os, tempfile, pathlib, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pytest
import test_gpu_parity as T
import kmergma_jl_b200 as K
from oracle import oracle as O
K.default_context()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = []
for seed in range(lo, hi):
    with tempfile.TemporaryDirectory() as td:
        try:
            T.test_fuzz_vs_oracle.__wrapped__(K, O, pathlib.Path(td), seed) if hasattr(T.test_fuzz_vs_oracle, "__wrapped__") else T.test_fuzz_vs_oracle(K, O, pathlib.Path(td), seed)
        except pytest.skip.Exception:
            pass
        except Exception as e:
            bad.append(seed); print("SEED", seed, "FAILED:", repr(e)[:300], flush=True); traceback.print_exc(limit=3)
    if seed % 20 == 0:
        print("... up to seed", seed, "failures so far:", bad, flush=True)
print("done", lo, hi, "failures:", bad)
