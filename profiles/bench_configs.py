"""Device-resident step time of the OTHER BASELINE configs (bench.py times cfg2): cfg3 (2x3 grid, 2^22 slots,
10 proxy-leader partitions) and cfg5 (vanilla Mencius, 7 servers, 2^20 slots).  Same method as bench.py's
`value`: inputs resident in HBM, a fresh slot window per step (state touched per step > L2), CUDA events on the
engine's stream, W warm-up + K timed steps.
    python profiles/bench_configs.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frankenpaxos_b200 import VANILLA_MENCIUS, Engine  # noqa: E402
from frankenpaxos_b200 import traces as T  # noqa: E402

dev = torch.device("cuda", 0)
K, W = 4, 2
S = K + W


def td(x):
    return torch.from_numpy(x.view(np.int32).reshape(len(x), -1)).to(dev)


def timed(eng, step):
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    for s in range(W):
        step(s)
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ext)
    for k in range(K):
        step(W + k)
    e1.record(ext)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def cfg3():
    cfg, n = T.config_by_name("cfg3")
    q = 2
    eng = Engine(slot_capacity=S * n, max_batch=q * n, overflow_capacity=1 << 10, **cfg)
    a, p, b = T.workload(3, cfg, n, partitions=10)
    ins = []
    for s in range(S):
        def rb(rec):
            out = rec.copy(); out["slot"] = out["slot"] + s * n
            return td(out)
        ins.append((rb(a), rb(p), rb(b)))
    o_p2b = torch.empty((q * n, 4), dtype=torch.int32, device=dev); o_nack = torch.empty((q * n, 2), dtype=torch.int32, device=dev)
    o_ch = torch.empty((q * n, 2), dtype=torch.int32, device=dev); wm = torch.zeros(1, dtype=torch.int32, device=dev)

    def step(s):
        da, dp, db = ins[s]
        eng.proxyleader_arm_dev(da.data_ptr(), n)
        eng.acceptor_phase2a_dev(dp.data_ptr(), q * n, o_p2b.data_ptr(), o_nack.data_ptr())
        eng.proxyleader_phase2b_dev(db.data_ptr(), q * n, o_ch.data_ptr())
        eng.replica_chosen_last_dev(o_ch.data_ptr())
        eng.chosen_watermark_dev(wm.data_ptr())
    ms = timed(eng, step)
    r = eng.sync()
    assert r.n_chosen == n and r.n_nack == 0 and r.watermark == S * n, (r.n_chosen, r.watermark)
    eng.close()
    return {"config": "cfg3: flexible 2x3 grid, f=1, thrifty quorum = one column (Q=2), 2^22 slots per step, Phase2b "
                      "shuffled within 10 proxy-leader partitions (slot % 10)", "slots_per_step": n, "ms_per_step": ms,
            "slots_per_s": n / (ms * 1e-3), "algorithmic_GB/s": (56 * q + 24) * n / (ms * 1e-3) / 1e9,
            "calls": "arm + acceptor_phase2a + proxyleader_phase2b + replica_chosen + watermark"}


def cfg5():
    cfg, n = T.config_by_name("cfg5")
    f, srv = cfg["f"], cfg["acceptors_per_group"]
    eng = Engine(slot_capacity=S * n, max_batch=(srv - 1) * n, protocol=VANILLA_MENCIUS, **cfg)
    req, p, b = T.vanilla_cfg5(5, f, n)
    ins = []
    for s in range(S):
        def rb(rec, fix_dst=False):
            out = rec.copy(); out["slot"] = out["slot"] + s * n
            return out
        r2 = rb(req); r2["dst"] = r2["slot"] % srv                      # the owner moves with the window offset
        shift = (s * n) % srv
        p2 = rb(p); p2["dst"] = (p2["dst"] + shift) % srv               # every server index rotates with the owner
        b2 = rb(b); b2["acceptor"] = (b2["acceptor"] + shift) % srv
        ins.append((td(r2), td(p2), td(b2)))
    nrec = (srv - 1) * n
    o_rep = torch.empty((nrec, 4), dtype=torch.int32, device=dev); o_ch = torch.empty((nrec, 2), dtype=torch.int32, device=dev)

    def step(s):
        dr, dp, db = ins[s]
        eng.vm_client_request_dev(dr.data_ptr(), n)
        eng.vm_phase2a_dev(dp.data_ptr(), nrec, o_rep.data_ptr())
        eng.proxyleader_phase2b_dev(db.data_ptr(), nrec, o_ch.data_ptr())
    ms = timed(eng, step)
    r = eng.sync()
    assert r.status == 0 and r.n_chosen == n, (r.status, r.n_chosen)
    # per-kernel split of one more step (CUDA events between the calls)
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    eng.reset()
    dr, dp, db = ins[0]
    ev[0].record(ext); eng.vm_client_request_dev(dr.data_ptr(), n)
    ev[1].record(ext); eng.vm_phase2a_dev(dp.data_ptr(), nrec, o_rep.data_ptr())
    ev[2].record(ext); eng.proxyleader_phase2b_dev(db.data_ptr(), nrec, o_ch.data_ptr())
    ev[3].record(ext)
    eng.sync()
    split = {"vm_client_request_us": ev[0].elapsed_time(ev[1]) * 1e3, "vm_phase2a_us": ev[1].elapsed_time(ev[2]) * 1e3,
             "tally_us": ev[2].elapsed_time(ev[3]) * 1e3}
    eng.close()
    return {"split": split, "config": "cfg5: vanilla Mencius n=7 f=3, owner = slot % 7, 2^20 slots per step, 6 Phase2a + 6 Phase2b per "
                      "slot, both shuffled", "slots_per_step": n, "ms_per_step": ms, "slots_per_s": n / (ms * 1e-3),
            "messages_per_s": (1 + 2 * (srv - 1)) * n / (ms * 1e-3),
            "calls": "vm_client_request (arm + own vote) + vm_phase2a + proxyleader_phase2b"}


only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""     # (FPX_LIB_OVERRIDE=... --only cfg5: A/B of a tuning variant)
res = {"method": f"{W} warm-up + {K} timed steps, CUDA events on the engine's stream", "lib": os.environ.get("FPX_LIB_OVERRIDE", "default")}
if only in ("", "cfg3"):
    res["cfg3"] = cfg3()
if only in ("", "cfg5"):
    res["cfg5"] = cfg5()
print(json.dumps(res))
