import sys, json, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import kmergma_jl_b200 as K
from oracle import oracle as O
from conftest import TF, GENOME
RUN_DT = np.dtype([("record", "<i4"), ("profile", "<i4"), ("t_first", "<i8"), ("t_last", "<i8"), ("t_argmin", "<i8"),
                   ("D_min", "<i8"), ("flags", "<u4"), ("reserved", "<u4")])
rvs, wss, cons, inv = K.cluster_ref_API(TF, 6)
rvs, wss, cons = K.eliminate_null_params(rvs, wss, cons, inv)
out = {}
for name, thrs in (("t45", [45]*6), ("t3", [20, 50, 30, 44, 25, 41])):
    for dense in (False, True):
        res = []
        o = K.Omn_KmerGMA(genome_path=GENOME, refVecs=rvs, windowsizes=wss, consensus_seqs=cons, resultVec=res, thr_vec=thrs, buff=0, align_hits=False, dense=dense)
        runs = np.frombuffer(o.runs.tobytes(), dtype=RUN_DT)
        out[f"{name}_dense{int(dense)}"] = {"hits": [(h.record, h.profile, h.cmi, h.first, h.last, h.D, h.dist, h.flags) for h in o.hits],
                                            "runs": [tuple(int(x) for x in r) for r in runs.tolist()]}
    for sc in (1.0, 1 - 2e-9, 1 + 2e-9):
        oh, _, _ = O.Omn_KmerGMA(GENOME, [np.asarray(v) for v in rvs], wss, cons, thr_vec=[t * sc for t in thrs], buff=0, align_hits=False)
        out[f"{name}_oracle_{sc!r}"] = [(h.record, h.kfv, h.cmi, h.first, h.last, h.dist) for h in oh]
out["members"] = [int(v.n_refs) for v in rvs]
json.dump(out, open('/root/repo/gpurun_out/dbg_cluster.json', 'w'))
print("ok")
