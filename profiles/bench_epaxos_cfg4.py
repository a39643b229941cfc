"""BASELINE cfg4 (EPaxos, 5 replicas, 20 % conflicts, 2^20 instances) as seen by replica 0:
device time of each handler batch (CUDA events around the kernels, copies excluded).
    python profiles/bench_epaxos_cfg4.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frankenpaxos_b200 import traces as T  # noqa: E402
from frankenpaxos_b200.epaxos import EpaxosReplica  # noqa: E402

f, n, N = 2, 5, 1 << 20
lead, pa, ok = T.epaxos_cfg4(0, f=f, n_instances=N, me=0)
eng = EpaxosReplica(f, 0, N // n + 2, max_batch=1 << 20)
res = {}
eng.lead(lead); res["lead (transitionToPreAcceptPhase)"] = (len(lead), eng.last_kernel_ms())
rep = eng.preaccept(pa); res["preaccept (handlePreAccept + dep union)"] = (len(pa), eng.last_kernel_ms())
ev = eng.preacceptok(ok); res["preacceptok (tally + fast-path vote + slow-path union)"] = (len(ok), eng.last_kernel_ms())
out = {"config": "cfg4: EPaxos n=5 f=2, 2^20 instances, BernoulliSingleKeyWorkload(0.2), replica 0's view",
       "fast_commits": int((ev[:, 0] == 1).sum()), "slow_paths": int((ev[:, 0] == 2).sum()),
       "calls": {k: {"messages": m, "kernel_ms": ms, "messages_per_s": m / (ms * 1e-3)} for k, (m, ms) in res.items()}}
print(json.dumps(out))
