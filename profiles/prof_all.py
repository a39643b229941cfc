"""One pass over every kernel family of libfpx.so at bench-like sizes, for `ncu --set full` (profiles/README.md):
cfg2 steps through fpx_step_dev (acceptor, arm, fused tally), the stand-alone replica / watermark kernels, one
cfg5 step (vanilla Mencius kernels, tally<16>), the EPaxos handlers (cfg4 shape), the wire codec, the conflict
index and the dependency graph.    ncu ... python profiles/prof_all.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from frankenpaxos_b200 import VANILLA_MENCIUS, Engine  # noqa: E402
from frankenpaxos_b200 import traces as T  # noqa: E402
from frankenpaxos_b200.epaxos import ConflictIndex, DependencyGraph, EpaxosReplica  # noqa: E402

dev = torch.device("cuda", 0)
td = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.int32).reshape(len(x), -1)).to(dev)

# ---- cfg2: three fused steps, then the separate replica / watermark kernels on a fourth window
cfg, n = bench.CFG, 1 << 20
eng = Engine(slot_capacity=5 * n, max_batch=3 * n, overflow_capacity=1 << 10, **cfg)
o1 = torch.empty((3 * n, 4), dtype=torch.int32, device=dev); o2 = torch.empty((3 * n, 2), dtype=torch.int32, device=dev)
o3 = torch.empty((3 * n, 2), dtype=torch.int32, device=dev); wm = torch.zeros(1, dtype=torch.int32, device=dev)
for s in range(4):
    a, p, b = T.workload(s, cfg, n, slot0=s * n)
    da, dp, db = td(a), td(p), td(b)
    if s < 3:
        eng.step_dev(da.data_ptr(), n, dp.data_ptr(), 3 * n, o1.data_ptr(), o2.data_ptr(), db.data_ptr(), 3 * n, o3.data_ptr(), wm.data_ptr())
    else:
        eng.proxyleader_arm_dev(da.data_ptr(), n)
        eng.acceptor_phase2a_dev(dp.data_ptr(), 3 * n, o1.data_ptr(), o2.data_ptr())
        eng.proxyleader_phase2b_dev(db.data_ptr(), 3 * n, o3.data_ptr())
        eng.replica_chosen_last_dev(o3.data_ptr())
        eng.chosen_watermark_dev(wm.data_ptr())
    r = eng.sync()
    assert r.n_chosen == n and r.watermark == (s + 1) * n
# ---- wire codec on the last step's votes
nb = len(b)
cap = 46 * nb
d_bytes = torch.empty(cap + 64, dtype=torch.uint8, device=dev); d_offs = torch.empty(nb + 1, dtype=torch.int32, device=dev)
d_kind = torch.empty(nb, dtype=torch.int32, device=dev); d_out = torch.empty((nb, 4), dtype=torch.int32, device=dev)
eng.wire_encode_phase2b_dev(db.data_ptr(), nb, d_bytes.data_ptr(), cap, d_offs.data_ptr())
eng.wire_decode_inbound_dev(0, d_bytes.data_ptr(), d_offs.data_ptr(), nb, d_kind.data_ptr(), d_out.data_ptr())
eng.sync()
eng.close()
del o1, o2, o3, da, dp, db
# ---- cfg5: one vanilla Mencius step
c5, n5 = T.config_by_name("cfg5")
srv = c5["acceptors_per_group"]
e5 = Engine(slot_capacity=n5, max_batch=(srv - 1) * n5, protocol=VANILLA_MENCIUS, **c5)
req, p, b = T.vanilla_cfg5(5, c5["f"], n5)
dr, dp, db = td(req), td(p), td(b)
orep = torch.empty((len(p), 4), dtype=torch.int32, device=dev); och = torch.empty((len(b), 2), dtype=torch.int32, device=dev)
e5.vm_client_request_dev(dr.data_ptr(), n5)
e5.vm_phase2a_dev(dp.data_ptr(), len(p), orep.data_ptr())
e5.proxyleader_phase2b_dev(db.data_ptr(), len(b), och.data_ptr())
assert e5.sync().n_chosen == n5
e5.close()
del dr, dp, db, orep, och
# ---- EPaxos handlers, cfg4 shape
f, N = 2, 1 << 20
nrep = 2 * f + 1
lead, pa, ok = T.epaxos_cfg4(0, f=f, n_instances=N, me=0)
tdi = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(dev)
d_lead, d_pa, d_ok = tdi(lead), tdi(pa), tdi(ok)
d_rep = torch.zeros((len(pa), 4 + nrep), dtype=torch.int32, device=dev); d_ev = torch.zeros((len(ok), 2 + nrep), dtype=torch.int32, device=dev)
ep = EpaxosReplica(f, 0, N // nrep + 2, max_batch=1 << 20)
ep.lead_dev(d_lead.data_ptr(), len(lead)); ep.preaccept_dev(d_pa.data_ptr(), len(pa), d_rep.data_ptr())
ep.preacceptok_dev(d_ok.data_ptr(), len(ok), d_ev.data_ptr()); ep.sync()
ep.close()
# ---- conflict index (hot keys: BernoulliSingleKeyWorkload) and dependency graph
g = T.rng(3)
M = 1 << 17
ci = ConflictIndex(5, key_capacity=1 << 12, max_commands=M, max_keys=M)
is_set = g.random(M) < 0.2
ci.batch(np.arange(M) % 5, np.arange(M) // 5, is_set, [[0] if s else [1] for s in is_set])
ci.close()
V = 1 << 15
dg = DependencyGraph(key_capacity=V, dep_pool_capacity=4 * V, max_batch=V)
keys = np.arange(V)
dg.commit(keys, (keys * 7) % 101, [[int(k) - 1] if k % 64 else ([int(k) + 63] if k + 63 < V else []) for k in keys])
comps, _ = dg.execute_by_component()
assert sum(len(c) for c in comps) > 0
dg.close()
print("prof_all ok")
