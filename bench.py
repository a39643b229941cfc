#!/usr/bin/env python3
"""bench.py -- committed slots/sec of the quorum-vote hot path on B200.

Workload (BASELINE.json configs[1], "cfg2"): MultiPaxos, 5 acceptors (f=2),
thrifty quorum of 3, 2^20 slots in flight per GPU per step.  One STEP is one
pass of the hot path over one window of 2^20 fresh slots:
    arm 2^20 (slot, round)  ->  3*2^20 Phase2a at the acceptors (ballot CAS +
    vote cells + Phase2b stream)  ->  3*2^20 shuffled Phase2b at the proxy leader
    (tally + quorum check + ordered Chosen stream)  ->  replica log + chosen
    watermark (+ NCCL all-gather of the per-GPU watermark when N > 1).
`value` times K steps with every input already resident in HBM (distinct
buffers per step, > L2 in total, fresh state region per step);
`e2e` times the same K steps through the host-pointer C ABI from pinned host
buffers (H2D of every input, D2H of the Phase2b and Chosen replies).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SLOTS_PER_STEP = 1 << 20
CFG = dict(f=2, num_acceptor_groups=1, acceptors_per_group=5, flexible=False, num_leaders=3, num_replicas=3)
Q = CFG["f"] + 1
# Algorithmic bytes per committed slot (SURVEY.md 8(d) / DESIGN.md):
#   acceptor kernel  40*Q  = read 16Q (Phase2a) + write 8Q (vote cell) + write 16Q (Phase2b)
#   tally kernels    16Q+24 = read 16Q (Phase2b) + 8+8 slot state RMW + write 8 (Chosen)
B_ACCEPTOR = 40 * Q
B_TALLY = 16 * Q + 24
B_SLOT = B_ACCEPTOR + B_TALLY
N_BASE = 4  # distinct base traces; step s uses base s % N_BASE re-based onto its own slot window


def config_dict(n_gpus):
    """`config` of the JSON line: identical for the GPU arm and the reference arm."""
    nrec = Q * SLOTS_PER_STEP
    return {"workload": "cfg2: MultiPaxos f=2, 5 acceptors, thrifty quorum 3, 2^20 slots in flight "
                        "per GPU per step, Phase2b globally shuffled",
            "slots_per_step_per_gpu": SLOTS_PER_STEP, "records_per_step_per_gpu": 2 * nrec + SLOTS_PER_STEP,
            "sharding": f"slot % {n_gpus}", "l2": "distinct input buffers and a fresh state window every "
                                                 "step (inputs+state touched per step 196 MB > L2)",
            "bytes_per_slot_algorithmic": B_SLOT}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-slots", type=int, default=SLOTS_PER_STEP)
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the
    committed ncu --set full capture (profiles/r1_traffic.json); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            k = json.load(f)[kernel]
        return int(k["dram_bytes_read"] + k["dram_bytes_write"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------- CPU arms (oracle port)
_CPU_TRACES = {}


def cpu_run(sample_slots, threads):
    """The reference's path restated on the CPU (oracle/fpx_oracle.cc, std::map /
    std::set like the Scala collections), `threads` proxy-leader/acceptor
    partitions by slot % threads (the reference's own scale-out), one pass over a
    bounded sample of the cfg2 workload.  Returns slots/s."""
    from frankenpaxos_b200 import traces as T
    from oracle import fpx_oracle_py as O
    key = (sample_slots, threads)
    if key not in _CPU_TRACES:      # the trace is generated once; only the handlers are timed
        a, p, b = T.workload(12345, CFG, sample_slots)
        _CPU_TRACES[key] = [(a[a["slot"] % threads == t], p[p["slot"] % threads == t], b[b["slot"] % threads == t])
                            for t in range(threads)]
    parts = _CPU_TRACES[key]
    oras = [O.MultiPaxos(CFG["f"], 1, 5, False, 3, 3) for _ in range(threads)]
    done = [0] * threads

    def work(t):
        o, (aa, pp, bb) = oras[t], parts[t]
        o.arm(aa)
        st, _, pb, nk = o.acceptor_phase2a(pp)
        st, _, c = o.proxyleader_phase2b(bb)
        o.replica_chosen(c)
        done[t] = len(c)

    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        [x.start() for x in th]
        [x.join() for x in th]
    dt = time.perf_counter() - t0
    assert sum(done) == sample_slots
    return sample_slots / dt, dt


def reference_arm(args, rank):
    """--impl reference: the reference's own CPU path.  The Scala/JVM reference
    cannot run (no JVM on the box, no offline build), so this is the C++ oracle
    PORT of the same handlers with the reference's data structures, on all host
    cores via the reference's own partitioning (slot % P)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    sample = args.cpu_sample_slots
    vals = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_run(min(sample, 1 << 15), threads)
    t_total = 0.0
    for _ in range(args.steps):
        v, dt = cpu_run(sample, threads)
        vals.append(v)
        t_total += dt
        if t_total > 120:
            break
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "committed slots/sec (simulated) at 1M in-flight slots",
        "value": value, "unit": "slots/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": "slots/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} slots x {len(vals)} passes, slot % {threads} partitions, "
                                   "C++ oracle port (JVM reference not runnable offline)"},
        "e2e": {"value": value, "unit": "slots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Polls NVML while the timed region runs (it is far shorter than nvidia-smi's
    sampling period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.active = index, [], False, False
        self.max_mhz, self.ok = 0, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample_now(self):
        """One sample.  (Sampling from the launch loop itself was tried: an NVML query costs ~0.1 ms of
        host time and showed up 1:1 in the step time, so the polling thread stays the only sampler; when
        it is starved of the GIL during the short timed region, summary() falls back to a wider window.)"""
        if not self.ok:
            return
        nv = self.nv
        try:
            mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.samples.append((time.perf_counter(), mhz, reasons))
        except Exception:
            pass

    def run(self):
        while self.ok and not self.stop_flag:
            self.sample_now()
            time.sleep(0.0001)

    def summary(self, windows):
        """windows: [(name, t0, t1)] in preference order; the first one holding >= 3 samples is used."""
        chosen, name = [], None
        for nm, t0, t1 in windows:
            chosen = [(m, r) for (t, m, r) in self.samples if t0 <= t <= t1]
            name = nm
            if len(chosen) >= 3:
                break
        if not self.ok or not chosen:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz or None, "reasons": ["unsampled"]}
        self_samples = chosen
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        seen = set()
        for _, r in self_samples:
            for bit, nm in names.items():
                if r & bit:
                    seen.add(nm)
        mhz = sorted(m for m, _ in self_samples)
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(seen),
                "samples": len(mhz), "window": name}


# --------------------------------------------------------------------------- our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from frankenpaxos_b200 import P2A, P2B, CHOSEN, NACK, Engine
    from frankenpaxos_b200 import traces as T

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product has no CPU path")
    N = world
    if args.gpus != N and world == 1 and args.gpus > 1:
        sys.exit("bench.py: launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if N > 1:
        dist.init_process_group("nccl", device_id=dev)

    K, W = args.steps, max(args.warmup, 3)
    S = K + W
    KE = min(K, 10)          # e2e steps: PCIe-bound and ~15x longer each, a bounded sample keeps the run short
    SE = KE + W
    total_windows = S + (0 if args.no_e2e else SE)
    if total_windows * SLOTS_PER_STEP * N >= (1 << 31):
        sys.exit(f"bench.py: (steps+warmup)*2^20*N must stay below 2^31 slots (int32 slot numbers)")
    n_slots_local = total_windows * SLOTS_PER_STEP
    eng = Engine(slot_capacity=n_slots_local * N, max_batch=Q * SLOTS_PER_STEP, overflow_capacity=1 << 10,
                 device=local_rank, shard_index=rank, shard_count=N, **CFG)
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)

    # ---- traces: N_BASE distinct seeded base traces on window 0; step s re-bases onto window s
    base = [T.workload(1000 * rank + b, CFG, SLOTS_PER_STEP) for b in range(N_BASE)]

    def rebase(rec, field, window):
        out = rec.copy()
        local = out[field].astype(np.int64) + window * SLOTS_PER_STEP
        out[field] = (local * N + rank).astype(np.int32)
        return out

    def step_inputs(window):
        a, p, b = base[window % N_BASE]
        return rebase(a, "slot", window), rebase(p, "slot", window), rebase(b, "slot", window)

    def to_dev(rec):
        t = torch.from_numpy(rec.view(np.int32).reshape(len(rec), -1))
        return t.to(dev, non_blocking=False)

    d_arm, d_p2a, d_p2b = [], [], []
    for s in range(S):
        a, p, b = step_inputs(s)
        d_arm.append(to_dev(a)); d_p2a.append(to_dev(p)); d_p2b.append(to_dev(b))
    nrec = Q * SLOTS_PER_STEP
    d_out_p2b = torch.empty((nrec, 4), dtype=torch.int32, device=dev)
    d_out_nack = torch.empty((nrec, 2), dtype=torch.int32, device=dev)
    d_out_chosen = torch.empty((nrec, 2), dtype=torch.int32, device=dev)
    # the per-step exchange (first-unchosen slot of every rank) runs on its own stream, double-buffered,
    # so that it overlaps the next step's kernels instead of sitting on the engine's stream
    d_wm = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(2)]
    d_wm_all = [torch.zeros(N, dtype=torch.int32, device=dev) for _ in range(2)]
    comm = torch.cuda.Stream(device=dev) if N > 1 else None
    ev_wm = [torch.cuda.Event() for _ in range(2)]
    ev_gathered = [None, None]
    torch.cuda.synchronize()

    def step(s, ev=None):
        eng.proxyleader_arm_dev(d_arm[s].data_ptr(), SLOTS_PER_STEP)
        if ev: ev[0].record(ext)
        eng.acceptor_phase2a_dev(d_p2a[s].data_ptr(), nrec, d_out_p2b.data_ptr(), d_out_nack.data_ptr())
        if ev: ev[1].record(ext)
        eng.proxyleader_phase2b_dev(d_p2b[s].data_ptr(), nrec, d_out_chosen.data_ptr())
        if ev: ev[2].record(ext)
        eng.replica_chosen_last_dev(d_out_chosen.data_ptr())
        par = s & 1
        if ev_gathered[par] is not None:
            ext.wait_event(ev_gathered[par])          # the all-gather of step s-2 has read this buffer
        eng.chosen_watermark_dev(d_wm[par].data_ptr())
        if N > 1:
            ev_wm[par].record(ext)
            comm.wait_event(ev_wm[par])
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(d_wm_all[par], d_wm[par])
                ev_gathered[par] = torch.cuda.Event()
                ev_gathered[par].record(comm)

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(W):
        step(s)
    r = eng.sync()
    assert r.n_chosen == SLOTS_PER_STEP and r.n_nack == 0, (r.n_chosen, r.n_nack)

    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launch_count
    barrier()
    t_timed0 = time.perf_counter()
    e_begin.record(ext)
    for k in range(K):
        step(W + k, evs[k])
    if N > 1:
        ext.wait_stream(comm)                         # the last exchanges are part of the timed region
    e_end.record(ext)
    barrier()
    t_timed1 = time.perf_counter()
    launches = eng.launch_count - launches0
    ms = e_begin.elapsed_time(e_end)
    r = eng.sync()
    assert r.status == 0 and r.n_chosen == SLOTS_PER_STEP and r.n_nack == 0
    exp_wm = (S * SLOTS_PER_STEP) * N + rank
    assert r.watermark == exp_wm, (r.watermark, exp_wm)
    if N > 1:   # the gathered frontiers of the last step: every rank's first-unchosen slot, global prefix = min
        last = d_wm_all[(S - 1) & 1].cpu().numpy()
        assert last.tolist() == [(S * SLOTS_PER_STEP) * N + g for g in range(N)], last

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if N > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = N * K * SLOTS_PER_STEP / (ms_max * 1e-3)

    acc_ms = float(np.mean([evs[k][0].elapsed_time(evs[k][1]) for k in range(K)]))
    tally_ms = float(np.mean([evs[k][1].elapsed_time(evs[k][2]) for k in range(K)]))
    peak, peak_src = peaks()
    acc_gbs = B_ACCEPTOR * SLOTS_PER_STEP / (acc_ms * 1e-3) / 1e9
    tally_gbs = B_TALLY * SLOTS_PER_STEP / (tally_ms * 1e-3) / 1e9

    # ---- e2e: host-pointer C ABI, pinned host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        h_in = []
        for s in range(SE):
            a, p, b = step_inputs(S + s)
            h_in.append(tuple(torch.from_numpy(x.view(np.int32).reshape(len(x), -1)).pin_memory() for x in (a, p, b)))
        h_p2b = torch.empty((nrec, 4), dtype=torch.int32).pin_memory()
        h_nack = torch.empty((nrec, 2), dtype=torch.int32).pin_memory()
        h_chosen = torch.empty((nrec, 2), dtype=torch.int32).pin_memory()
        L = eng._L
        n1, n2, n3, err = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
        wm = ctypes.c_int32()

        def e2e_step(s):
            a, p, b = h_in[s]
            st = L.fpx_proxyleader_arm(eng.h, a.data_ptr(), SLOTS_PER_STEP, ctypes.byref(err))
            st |= L.fpx_acceptor_phase2a(eng.h, p.data_ptr(), nrec, h_p2b.data_ptr(), ctypes.byref(n1),
                                         h_nack.data_ptr(), ctypes.byref(n2), ctypes.byref(err))
            st |= L.fpx_proxyleader_phase2b(eng.h, b.data_ptr(), nrec, h_chosen.data_ptr(), ctypes.byref(n3),
                                            ctypes.byref(err))
            st |= L.fpx_replica_chosen(eng.h, h_chosen.data_ptr(), n3.value, ctypes.byref(err))
            st |= L.fpx_chosen_watermark(eng.h, ctypes.byref(wm))
            assert st == 0 and n3.value == SLOTS_PER_STEP and n1.value == nrec, (st, n1.value, n3.value)
            if N > 1:
                d_wm[0].fill_(wm.value)
                dist.all_gather_into_tensor(d_wm_all[0], d_wm[0])

        for s in range(W):
            e2e_step(s)
        barrier()
        t0 = time.perf_counter()
        for k in range(KE):
            e2e_step(W + k)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if N > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": N * KE * SLOTS_PER_STEP / float(t.item()), "unit": "slots/s", "steps": KE,
               "h2d_bytes_per_step": 16 * SLOTS_PER_STEP + 2 * 16 * nrec + 8 * SLOTS_PER_STEP,
               "d2h_bytes_per_step": 16 * nrec + 8 * SLOTS_PER_STEP + 4,
               "ms_per_step": 1e3 * float(t.item()) / KE,
               "api": "fpx_proxyleader_arm + fpx_acceptor_phase2a + fpx_proxyleader_phase2b + "
                      "fpx_replica_chosen + fpx_chosen_watermark (host pointers, pinned)"}
    sampler.stop_flag = True
    t_all1 = time.perf_counter()
    clocks = sampler.summary([("timed region", t_timed0, t_timed1),
                              ("timed region + e2e region (timed region too short to sample 3 times)", t_timed0, t_all1)])

    cpu = None
    if rank == 0 and N == 1:
        v, dt = cpu_run(args.cpu_sample_slots, 1)
        cpu = {"value": v, "unit": "slots/s", "cores": 1, "kind": "port",
               "sample": f"{args.cpu_sample_slots} slots of the cfg2 workload, one pass ({dt:.1f} s), "
                         "single-threaded C++ oracle port of the Scala handlers (JVM reference not runnable offline)"}

    if rank == 0:
        line = {
            "metric": "committed slots/sec (simulated) at 1M in-flight slots",
            "value": value, "unit": "slots/s", "n_gpus": N, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": config_dict(N),
            # the dominant kernel by time (43 % of the step in profiles/r1_launches_final.csv) is the tally
            "roofline": {"bound": "hbm", "kernel": "tally_kernel", "achieved": tally_gbs, "peak": peak,
                         "unit": "GB/s", "frac": tally_gbs / peak, "traffic": ncu_traffic("tally_kernel"),
                         "algorithmic_bytes_per_launch": B_TALLY * SLOTS_PER_STEP, "ms_per_launch": tally_ms,
                         "peak_source": peak_src,
                         "note": "shuffled Phase2b stream: bound by L1TEX wavefronts of the 2 divergent row accesses "
                                 "per vote, not by DRAM bytes (DESIGN.md section 4)"},
            "kernels": {"acceptor_phase2a_kernel": {"ms": acc_ms, "GB/s": acc_gbs, "frac": acc_gbs / peak,
                                                    "algorithmic_bytes_per_launch": B_ACCEPTOR * SLOTS_PER_STEP,
                                                    "traffic": ncu_traffic("acceptor_phase2a_kernel")},
                        "tally_kernel": {"ms": tally_ms, "GB/s": tally_gbs, "frac": tally_gbs / peak,
                                         "algorithmic_bytes_per_launch": B_TALLY * SLOTS_PER_STEP,
                                         "traffic": ncu_traffic("tally_kernel")},
                        "whole_step_GB/s": B_SLOT * SLOTS_PER_STEP / (ms_max / K * 1e-3) / 1e9},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if N > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
