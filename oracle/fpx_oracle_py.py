"""ctypes binding of the CPU oracle (oracle/fpx_oracle.cc).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfpx_oracle.so")

P2A = np.dtype([("slot", "<i4"), ("round", "<i4"), ("value_id", "<i4"), ("dst", "<i4")])
P2B = np.dtype([("group", "<i4"), ("acceptor", "<i4"), ("slot", "<i4"), ("round", "<i4")])
CHOSEN = np.dtype([("slot", "<i4"), ("value_id", "<i4")])
NACK = np.dtype([("leader", "<i4"), ("round", "<i4")])
P2A_RANGE = np.dtype([("slot_start", "<i4"), ("slot_end", "<i4"), ("round", "<i4"), ("dst", "<i4")])
P2B_RANGE = np.dtype([("dst", "<i4"), ("slot_start", "<i4"), ("slot_end", "<i4"), ("round", "<i4")])
CHOSEN_RANGE = np.dtype([("slot_start", "<i4"), ("slot_end", "<i4")])
VM_SKIP = np.dtype([("server", "<i4"), ("slot_start", "<i4"), ("slot_stop", "<i4"), ("own", "<i4")])
NOOP = -(1 << 31)


def build(force=False):
    src = os.path.join(_HERE, "fpx_oracle.cc")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libfpx_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, i64p = C.c_void_p, C.c_int, C.POINTER(C.c_int64)
        ip = C.POINTER(C.c_int)
        L.fpo_quorum_eval.argtypes = [i32, ip, i32, i32, i32, ip, i32]
        L.fpo_quorum_eval.restype = i32
        for name in ["fpo_ips_new"]:
            getattr(L, name).restype = vp
        L.fpo_ips_from_set.argtypes = [ip, i32]; L.fpo_ips_from_set.restype = vp
        L.fpo_ips_from_watermark_values.argtypes = [i32, ip, i32]; L.fpo_ips_from_watermark_values.restype = vp
        L.fpo_ips_free.argtypes = [vp]
        L.fpo_ips_clone.argtypes = [vp]; L.fpo_ips_clone.restype = vp
        for name in ["fpo_ips_add", "fpo_ips_contains"]:
            getattr(L, name).argtypes = [vp, i32]; getattr(L, name).restype = i32
        for name in ["fpo_ips_watermark", "fpo_ips_num_values", "fpo_ips_size"]:
            getattr(L, name).argtypes = [vp]; getattr(L, name).restype = i32
        for name in ["fpo_ips_values", "fpo_ips_materialize"]:
            getattr(L, name).argtypes = [vp, ip, i32]; getattr(L, name).restype = i32
        L.fpo_ips_equals.argtypes = [vp, vp]; L.fpo_ips_equals.restype = i32
        for name in ["fpo_ips_union", "fpo_ips_diff", "fpo_ips_diff_iterator"]:
            getattr(L, name).argtypes = [vp, vp]; getattr(L, name).restype = vp
        for name in ["fpo_ips_add_all", "fpo_ips_subtract_all"]:
            getattr(L, name).argtypes = [vp, vp]; getattr(L, name).restype = None
        L.fpo_ips_subtract_one.argtypes = [vp, i32]; L.fpo_ips_subtract_one.restype = None
        L.fpo_ips_diff_iterator_free.argtypes = [vp]
        L.fpo_ips_diff_iterator_has_next.argtypes = [vp]; L.fpo_ips_diff_iterator_has_next.restype = i32
        L.fpo_ips_diff_iterator_next.argtypes = [vp]; L.fpo_ips_diff_iterator_next.restype = i32
        L.fpo_topone_new.argtypes = [i32]; L.fpo_topone_new.restype = vp
        L.fpo_topone_free.argtypes = [vp]
        L.fpo_topone_put.argtypes = [vp, i32, i32]
        L.fpo_topone_merge.argtypes = [vp, vp]
        L.fpo_topone_get.argtypes = [vp, ip]
        L.fpo_dg_new.argtypes = []; L.fpo_dg_new.restype = vp
        L.fpo_dg_free.argtypes = [vp]
        L.fpo_dg_commit.argtypes = [vp, i32, i32, ip, i32]
        L.fpo_dg_update_executed.argtypes = [vp, ip, i32]
        L.fpo_dg_execute_by_component.argtypes = [vp, i32, ip, ip, i32, ip, ip]; L.fpo_dg_execute_by_component.restype = i32
        L.fpo_kvci_new.argtypes = [i32]; L.fpo_kvci_new.restype = vp
        L.fpo_kvci_free.argtypes = [vp]
        L.fpo_kvci_put.argtypes = [vp, i32, i32, i32, ip, i32]
        L.fpo_kvci_put_snapshot.argtypes = [vp, i32, i32]
        L.fpo_kvci_top_one_conflicts.argtypes = [vp, i32, ip, i32, ip]
        L.fpo_qw_new.argtypes = [i32]; L.fpo_qw_new.restype = vp
        L.fpo_qw_free.argtypes = [vp]
        L.fpo_qw_update.argtypes = [vp, i32, i32]
        L.fpo_qw_watermark.argtypes = [vp, i32]; L.fpo_qw_watermark.restype = i32
        L.fpo_bm_new.argtypes = [i32]; L.fpo_bm_new.restype = vp
        L.fpo_bm_free.argtypes = [vp]
        L.fpo_bm_get.argtypes = [vp, i32]; L.fpo_bm_get.restype = i32
        L.fpo_bm_put.argtypes = [vp, i32, i32]
        L.fpo_bm_gc.argtypes = [vp, i32]
        L.fpo_rr_leader.argtypes = [i32, i32]; L.fpo_rr_leader.restype = i32
        L.fpo_rr_next_classic_round.argtypes = [i32, i32, i32]; L.fpo_rr_next_classic_round.restype = i32
        L.fpo_mp_new.argtypes = [i32] * 6; L.fpo_mp_new.restype = vp
        L.fpo_mencius_new.argtypes = [i32] * 6; L.fpo_mencius_new.restype = vp
        L.fpo_mp_free.argtypes = [vp]
        L.fpo_mp_arm.argtypes = [vp, vp, i32, i64p]; L.fpo_mp_arm.restype = i32
        L.fpo_mp_acceptor_phase2a.argtypes = [vp, vp, i32, vp, ip, vp, ip, i64p]
        L.fpo_mp_acceptor_phase2a.restype = i32
        L.fpo_mp_proxyleader_phase2b.argtypes = [vp, vp, i32, vp, ip, i64p]
        L.fpo_mp_proxyleader_phase2b.restype = i32
        L.fpo_mp_replica_chosen.argtypes = [vp, vp, i32]; L.fpo_mp_replica_chosen.restype = i32
        L.fpo_mp_executed_watermark.argtypes = [vp]; L.fpo_mp_executed_watermark.restype = i32
        L.fpo_mp_first_hole.argtypes = [vp]; L.fpo_mp_first_hole.restype = i32
        L.fpo_mp_arm_range.argtypes = [vp, vp, i32, i32, i64p]; L.fpo_mp_arm_range.restype = i32
        L.fpo_mp_acceptor_noop_range.argtypes = [vp, vp, i32, i32, vp, ip, vp, ip, i64p]
        L.fpo_mp_acceptor_noop_range.restype = i32
        L.fpo_mp_range_phase2b.argtypes = [vp, vp, i32, vp, ip, i64p]; L.fpo_mp_range_phase2b.restype = i32
        L.fpo_mp_replica_chosen_range.argtypes = [vp, vp, i32, i32, i64p]; L.fpo_mp_replica_chosen_range.restype = i32
        L.fpo_mp_snapshot_acceptor.argtypes = [vp, i32, i32, ip, ip, i32, i32, vp, vp]
        L.fpo_mp_snapshot_log.argtypes = [vp, i32, i32, vp]
        L.fpo_mp_phase1a.argtypes = [vp, i32, i32, i32]; L.fpo_mp_phase1a.restype = i32
        L.fpo_mp_safe_values.argtypes = [vp, C.c_uint, i32, i32, vp, vp, ip]
        L.fpo_ep_new.argtypes = [i32, i32]; L.fpo_ep_new.restype = vp
        L.fpo_ep_free.argtypes = [vp]
        L.fpo_ep_lead.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, i32]; L.fpo_ep_lead.restype = i32
        L.fpo_ep_pre_accept.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, ip]
        L.fpo_ep_accept.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, ip]
        L.fpo_ep_pre_accept_ok.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, ip]
        L.fpo_ep_accept_ok.argtypes = [vp, i32, i32, i32, i32, i32, vp, ip]
        L.fpo_ep_entry.argtypes = [vp, i32, i32, vp]; L.fpo_ep_entry.restype = i32
        L.fpo_ep_leader_kind.argtypes = [vp, i32, i32]; L.fpo_ep_leader_kind.restype = i32
        L.fpo_ep_largest_ballot.argtypes = [vp, vp]
        L.fpo_wire_decode_inbound.argtypes = [i32, vp, vp, i32, i32, i32, vp, vp, i64p]; L.fpo_wire_decode_inbound.restype = i32
        L.fpo_wire_encode_phase2b.argtypes = [vp, i32, vp, vp]; L.fpo_wire_encode_phase2b.restype = C.c_int64
        L.fpo_wire_encode_nack.argtypes = [vp, i32, vp, vp]; L.fpo_wire_encode_nack.restype = C.c_int64
        L.fpo_wire_encode_mencius_phase2b.argtypes = [vp, i32, vp, vp]; L.fpo_wire_encode_mencius_phase2b.restype = C.c_int64
        L.fpo_wire_encode_chosen.argtypes = [vp, i32, vp, vp, i32, vp, vp, i64p]; L.fpo_wire_encode_chosen.restype = C.c_int64
        L.fpo_vm_new.argtypes = [i32]; L.fpo_vm_new.restype = vp
        L.fpo_vm_free.argtypes = [vp]
        L.fpo_vm_client_request.argtypes = [vp, vp, i32, i64p]; L.fpo_vm_client_request.restype = i32
        L.fpo_vm_phase2a.argtypes = [vp, vp, i32, vp, i64p]; L.fpo_vm_phase2a.restype = i32
        L.fpo_vm_phase2b.argtypes = [vp, vp, i32, vp, ip, i64p]; L.fpo_vm_phase2b.restype = i32
        L.fpo_vm_learn_chosen.argtypes = [vp, vp, i32]
        L.fpo_vm_skip.argtypes = [vp, vp, i32, i32, i64p]; L.fpo_vm_skip.restype = i32
        L.fpo_vm_snapshot.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        _lib = L
    return _lib


def _iarr(xs):
    xs = list(xs)
    return (C.c_int * max(1, len(xs)))(*xs), len(xs)


QUORUM_PREDS = {"isReadQuorum": 0, "isWriteQuorum": 1, "isSuperSetOfReadQuorum": 2,
                "isSuperSetOfWriteQuorum": 3}


def quorum_eval(kind, members, which, xs):
    """kind: 'grid' (members = list of rows) | 'simple_majority' | 'unanimous_writes'.
    Returns 0/1, or 2 where the reference `require` throws."""
    if kind == "grid":
        rows, cols = len(members), len(members[0])
        flat = [x for r in members for x in r]
        k = 0
    else:
        rows, cols = 1, len(members)
        flat = list(members)
        k = 1 if kind == "simple_majority" else 2
    m, _ = _iarr(flat)
    x, n = _iarr(xs)
    return lib().fpo_quorum_eval(k, m, rows, cols, QUORUM_PREDS[which] if isinstance(which, str) else which, x, n)


class IntPrefixSet:
    def __init__(self, handle=None):
        self.h = handle if handle is not None else lib().fpo_ips_new()

    @classmethod
    def from_set(cls, xs):
        a, n = _iarr(sorted(xs))
        return cls(lib().fpo_ips_from_set(a, n))

    @classmethod
    def from_watermark_values(cls, w, xs):
        a, n = _iarr(sorted(xs))
        return cls(lib().fpo_ips_from_watermark_values(w, a, n))

    def __del__(self):
        try:
            lib().fpo_ips_free(self.h)
        except Exception:
            pass

    def clone(self): return IntPrefixSet(lib().fpo_ips_clone(self.h))
    def add(self, x): return bool(lib().fpo_ips_add(self.h, x))
    def contains(self, x): return bool(lib().fpo_ips_contains(self.h, x))
    def watermark(self): return lib().fpo_ips_watermark(self.h)
    def size(self): return lib().fpo_ips_size(self.h)

    def values(self):
        n = lib().fpo_ips_num_values(self.h)
        buf = (C.c_int * max(1, n))()
        lib().fpo_ips_values(self.h, buf, n)
        return set(buf[:n])

    def materialize(self):
        n = lib().fpo_ips_size(self.h)
        buf = (C.c_int * max(1, n))()
        lib().fpo_ips_materialize(self.h, buf, n)
        return set(buf[:n])

    def __eq__(self, o): return bool(lib().fpo_ips_equals(self.h, o.h))
    def union(self, o): return IntPrefixSet(lib().fpo_ips_union(self.h, o.h))
    def diff(self, o): return IntPrefixSet(lib().fpo_ips_diff(self.h, o.h))
    def add_all(self, o): lib().fpo_ips_add_all(self.h, o.h); return self
    def subtract_all(self, o): lib().fpo_ips_subtract_all(self.h, o.h); return self
    def subtract_one(self, x): lib().fpo_ips_subtract_one(self.h, x); return self

    def diff_iterator(self, o):
        return DiffIterator(self, o)


class DiffIterator:
    def __init__(self, me, other):
        self._keep = (me, other)
        self.h = lib().fpo_ips_diff_iterator(me.h, other.h)

    def __del__(self):
        try:
            lib().fpo_ips_diff_iterator_free(self.h)
        except Exception:
            pass

    def has_next(self): return bool(lib().fpo_ips_diff_iterator_has_next(self.h))
    def next(self): return lib().fpo_ips_diff_iterator_next(self.h)


class MultiPaxos:
    """Sequential restatement of the multipaxos Acceptor / ProxyLeader / Replica
    handlers on the quorum-vote path; same call shapes as frankenpaxos_b200.Engine."""

    def __init__(self, f, groups, per_group, flexible=False, num_leaders=None, num_replicas=None,
                 mencius_leader_groups=0):
        num_leaders = f + 1 if num_leaders is None else num_leaders
        num_replicas = f + 1 if num_replicas is None else num_replicas
        if mencius_leader_groups:
            self.h = lib().fpo_mencius_new(f, mencius_leader_groups, groups, per_group, num_leaders, num_replicas)
        else:
            self.h = lib().fpo_mp_new(f, groups, per_group, int(flexible), num_leaders, num_replicas)
        if not self.h:
            raise ValueError("Config.checkValid failed")

    def __del__(self):
        try:
            if self.h:
                lib().fpo_mp_free(self.h)
        except Exception:
            pass

    def arm(self, p2a):
        p2a = np.ascontiguousarray(p2a, dtype=P2A)
        err = C.c_int64(-1)
        st = lib().fpo_mp_arm(self.h, p2a.ctypes.data, len(p2a), C.byref(err))
        return st, err.value

    def acceptor_phase2a(self, p2a):
        p2a = np.ascontiguousarray(p2a, dtype=P2A)
        n = len(p2a)
        out = np.zeros(max(n, 1), dtype=P2B)
        nack = np.zeros(max(n, 1), dtype=NACK)
        n1, n2, err = C.c_int(0), C.c_int(0), C.c_int64(-1)
        st = lib().fpo_mp_acceptor_phase2a(self.h, p2a.ctypes.data, n, out.ctypes.data, C.byref(n1),
                                           nack.ctypes.data, C.byref(n2), C.byref(err))
        return st, err.value, out[:n1.value].copy(), nack[:n2.value].copy()

    def proxyleader_phase2b(self, p2b):
        p2b = np.ascontiguousarray(p2b, dtype=P2B)
        n = len(p2b)
        out = np.zeros(max(n, 1), dtype=CHOSEN)
        n1, err = C.c_int(0), C.c_int64(-1)
        st = lib().fpo_mp_proxyleader_phase2b(self.h, p2b.ctypes.data, n, out.ctypes.data, C.byref(n1),
                                              C.byref(err))
        return st, err.value, out[:n1.value].copy()

    def replica_chosen(self, chosen):
        chosen = np.ascontiguousarray(chosen, dtype=CHOSEN)
        return lib().fpo_mp_replica_chosen(self.h, chosen.ctypes.data, len(chosen))

    def executed_watermark(self):
        return lib().fpo_mp_executed_watermark(self.h)

    def first_hole(self):
        """Where executeLog would stop if it ran now (what fpx_chosen_watermark reports)."""
        return lib().fpo_mp_first_hole(self.h)

    # -- S/mencius Phase2aNoopRange path; `capacity` = the engine's slot_capacity (range precondition)
    def arm_range(self, recs, capacity):
        recs = np.ascontiguousarray(recs, dtype=P2A_RANGE)
        err = C.c_int64(-1)
        st = lib().fpo_mp_arm_range(self.h, recs.ctypes.data, len(recs), capacity, C.byref(err))
        return st, err.value

    def acceptor_noop_range(self, recs, capacity):
        recs = np.ascontiguousarray(recs, dtype=P2A_RANGE)
        n = len(recs)
        out = np.zeros(max(n, 1), dtype=P2B_RANGE)
        nack = np.zeros(max(n, 1), dtype=NACK)
        n1, n2, err = C.c_int(0), C.c_int(0), C.c_int64(-1)
        st = lib().fpo_mp_acceptor_noop_range(self.h, recs.ctypes.data, n, capacity, out.ctypes.data, C.byref(n1),
                                              nack.ctypes.data, C.byref(n2), C.byref(err))
        return st, err.value, out[:n1.value].copy(), nack[:n2.value].copy()

    def range_phase2b(self, recs):
        recs = np.ascontiguousarray(recs, dtype=P2B_RANGE)
        n = len(recs)
        out = np.zeros(max(n, 1), dtype=CHOSEN_RANGE)
        n1, err = C.c_int(0), C.c_int64(-1)
        st = lib().fpo_mp_range_phase2b(self.h, recs.ctypes.data, n, out.ctypes.data, C.byref(n1), C.byref(err))
        return st, err.value, out[:n1.value].copy()

    def replica_chosen_range(self, recs, capacity):
        recs = np.ascontiguousarray(recs, dtype=CHOSEN_RANGE)
        err = C.c_int64(-1)
        st = lib().fpo_mp_replica_chosen_range(self.h, recs.ctypes.data, len(recs), capacity, C.byref(err))
        return st, err.value

    def snapshot_acceptor(self, group, acceptor, first_slot, n_slots):
        vr = np.zeros(max(n_slots, 1), dtype=np.int32)
        vv = np.zeros(max(n_slots, 1), dtype=np.int32)
        r, m = C.c_int(0), C.c_int(0)
        lib().fpo_mp_snapshot_acceptor(self.h, group, acceptor, C.byref(r), C.byref(m), first_slot, n_slots,
                                       vr.ctypes.data, vv.ctypes.data)
        return r.value, m.value, vr[:n_slots], vv[:n_slots]

    def phase1a(self, group, acceptor, round_):
        return lib().fpo_mp_phase1a(self.h, group, acceptor, round_)

    def safe_values(self, responders, first_slot, n_slots):
        vr = np.zeros(max(n_slots, 1), dtype=np.int32); vv = np.zeros(max(n_slots, 1), dtype=np.int32)
        mx = C.c_int(-1)
        lib().fpo_mp_safe_values(self.h, responders, first_slot, n_slots, vr.ctypes.data, vv.ctypes.data, C.byref(mx))
        return vr[:n_slots], vv[:n_slots], mx.value

    def snapshot_log(self, first_slot, n_slots):
        v = np.zeros(max(n_slots, 1), dtype=np.int32)
        lib().fpo_mp_snapshot_log(self.h, first_slot, n_slots, v.ctypes.data)
        return v[:n_slots]


class EPaxos:
    """Sequential restatement of one epaxos.Replica's handlers on the path (dense deps).
    Same row formats as frankenpaxos_b200.epaxos.EpaxosReplica, one message at a time."""

    def __init__(self, f, index):
        self.f, self.n, self.index = f, 2 * f + 1, index
        self.h = lib().fpo_ep_new(f, index)
        self.saw_sparse = False

    def __del__(self):
        try:
            lib().fpo_ep_free(self.h)
        except Exception:
            pass

    def lead(self, rows):
        n = self.n
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 8 + n)
        for i, r in enumerate(rows):
            d = np.ascontiguousarray(r[8:8 + n])
            if lib().fpo_ep_lead(self.h, int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]),
                                 d.ctypes.data, int(r[6])) != 0:
                return -13, i
        return 0, -1

    def preaccept(self, rows):
        n = self.n
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 6 + 2 * n)
        out = np.zeros((len(rows), 4 + n), dtype=np.int32)
        sp = C.c_int(0)
        for i, r in enumerate(rows):
            a = np.ascontiguousarray(r[6:6 + n]); b = np.ascontiguousarray(r[6 + n:6 + 2 * n])
            lib().fpo_ep_pre_accept(self.h, int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]),
                                    a.ctypes.data, b.ctypes.data, out[i].ctypes.data, C.byref(sp))
            self.saw_sparse |= bool(sp.value)
        return out

    def accept(self, rows):
        n = self.n
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 6 + n)
        out = np.zeros((len(rows), 4 + n), dtype=np.int32)
        sp = C.c_int(0)
        for i, r in enumerate(rows):
            a = np.ascontiguousarray(r[6:6 + n])
            lib().fpo_ep_accept(self.h, int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]),
                                a.ctypes.data, out[i].ctypes.data, C.byref(sp))
        return out

    def preacceptok(self, rows):
        n = self.n
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 6 + n)
        out = np.zeros((len(rows), 2 + n), dtype=np.int32)
        sp = C.c_int(0)
        for i, r in enumerate(rows):
            a = np.ascontiguousarray(r[6:6 + n])
            lib().fpo_ep_pre_accept_ok(self.h, int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]),
                                       a.ctypes.data, out[i].ctypes.data, C.byref(sp))
            self.saw_sparse |= bool(sp.value)
        return out

    def acceptok(self, rows):
        n = self.n
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 6)
        out = np.zeros((len(rows), 2 + n), dtype=np.int32)
        sp = C.c_int(0)
        for i, r in enumerate(rows):
            lib().fpo_ep_accept_ok(self.h, int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]),
                                   out[i].ctypes.data, C.byref(sp))
        return out

    def entry(self, rep, num):
        out = np.zeros(7 + self.n, dtype=np.int32)
        if not lib().fpo_ep_entry(self.h, rep, num, out.ctypes.data):
            out[:] = 0
            out[1:5] = -1
        elif out[0] == 4:
            out[1:5] = -1
        lb = np.zeros(2, dtype=np.int32)
        lib().fpo_ep_largest_ballot(self.h, lb.ctypes.data)
        return out, lib().fpo_ep_leader_kind(self.h, rep, num), lb


class VanillaMencius:
    """Sequential restatement of vanillamencius.Server's normal-case handlers, all n
    servers co-located; same call shapes as Engine(protocol=VANILLA_MENCIUS)."""

    def __init__(self, f):
        self.f, self.n = f, 2 * f + 1
        self.h = lib().fpo_vm_new(f)

    def __del__(self):
        try:
            lib().fpo_vm_free(self.h)
        except Exception:
            pass

    def client_request(self, recs):
        recs = np.ascontiguousarray(recs, dtype=P2A)
        err = C.c_int64(-1)
        return lib().fpo_vm_client_request(self.h, recs.ctypes.data, len(recs), C.byref(err)), err.value

    def phase2a(self, recs):
        recs = np.ascontiguousarray(recs, dtype=P2A)
        out = np.zeros(max(len(recs), 1), dtype=P2B)
        err = C.c_int64(-1)
        st = lib().fpo_vm_phase2a(self.h, recs.ctypes.data, len(recs), out.ctypes.data, C.byref(err))
        return st, err.value, out[:len(recs)]

    def phase2b(self, recs):
        recs = np.ascontiguousarray(recs, dtype=P2B)
        out = np.zeros(max(len(recs), 1), dtype=CHOSEN)
        n1, err = C.c_int(0), C.c_int64(-1)
        st = lib().fpo_vm_phase2b(self.h, recs.ctypes.data, len(recs), out.ctypes.data, C.byref(n1), C.byref(err))
        return st, err.value, out[:n1.value].copy()

    proxyleader_phase2b = phase2b

    def learn_chosen(self, recs):
        recs = np.ascontiguousarray(recs, dtype=P2B)
        lib().fpo_vm_learn_chosen(self.h, recs.ctypes.data, len(recs))

    def skip(self, recs, capacity):
        recs = np.ascontiguousarray(recs, dtype=VM_SKIP)
        err = C.c_int64(-1)
        return lib().fpo_vm_skip(self.h, recs.ctypes.data, len(recs), capacity, C.byref(err)), err.value

    def snapshot(self, server, first_slot, n_slots):
        k = np.zeros(max(n_slots, 1), np.int32); r = np.zeros(max(n_slots, 1), np.int32); v = np.zeros(max(n_slots, 1), np.int32)
        lib().fpo_vm_snapshot(self.h, server, first_slot, n_slots, k.ctypes.data, r.ctypes.data, v.ctypes.data)
        return k[:n_slots], r[:n_slots], v[:n_slots]


# --------------------------------------------------------------------------- wire codec
WIRE_REC = np.dtype([("a", "<i4"), ("b", "<i4"), ("c", "<i4"), ("d", "<i4")])
WIRE_PROXYLEADER_INBOUND, WIRE_ACCEPTOR_INBOUND = 0, 1
WIRE_MENCIUS_PROXYLEADER_INBOUND, WIRE_MENCIUS_ACCEPTOR_INBOUND = 2, 3


def pack_messages(msgs):
    """list of bytes -> (uint8 array, int32 offsets[n+1])"""
    offs = np.zeros(len(msgs) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(m) for m in msgs])
    buf = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if msgs else np.zeros(0, dtype=np.uint8)
    return buf, offs


def wire_decode_inbound(inbound, buf, offs, lgroups=1, agroups=1):
    buf = np.ascontiguousarray(buf, dtype=np.uint8); offs = np.ascontiguousarray(offs, dtype=np.int32)
    n = len(offs) - 1
    kind = np.zeros(max(n, 1), dtype=np.int32); out = np.zeros(max(n, 1), dtype=WIRE_REC)
    err = C.c_int64(-1)
    pad = buf if len(buf) else np.zeros(1, dtype=np.uint8)
    st = lib().fpo_wire_decode_inbound(inbound, pad.ctypes.data, offs.ctypes.data, n, lgroups, agroups, kind.ctypes.data,
                                       out.ctypes.data, C.byref(err))
    return st, err.value, kind[:n], out[:n]


def _wire_encode(fn, recs, dtype, max_per):
    recs = np.ascontiguousarray(recs, dtype=dtype)
    n = len(recs)
    out = np.zeros(max(1, n * max_per), dtype=np.uint8); offs = np.zeros(n + 1, dtype=np.int32)
    total = fn(recs.ctypes.data, n, out.ctypes.data, offs.ctypes.data)
    return out[:total].copy(), offs


def wire_encode_phase2b(recs):
    return _wire_encode(lib().fpo_wire_encode_phase2b, recs, P2B, 46)


def wire_encode_mencius_phase2b(recs):
    return _wire_encode(lib().fpo_wire_encode_mencius_phase2b, recs, P2B, 46)


def wire_encode_nack(recs):
    return _wire_encode(lib().fpo_wire_encode_nack, recs, NACK, 13)


def wire_encode_chosen(recs, arena, value_offsets):
    recs = np.ascontiguousarray(recs, dtype=CHOSEN)
    arena = np.ascontiguousarray(arena, dtype=np.uint8); value_offsets = np.ascontiguousarray(value_offsets, dtype=np.int32)
    n = len(recs)
    lens = np.diff(value_offsets)
    cap = int(n * 24 + (lens[np.clip(recs["value_id"], 0, len(lens) - 1)].sum() if n and len(lens) else 0))
    out = np.zeros(max(1, cap), dtype=np.uint8); offs = np.zeros(n + 1, dtype=np.int32)
    err = C.c_int64(-1)
    pad = arena if len(arena) else np.zeros(1, dtype=np.uint8)
    total = lib().fpo_wire_encode_chosen(recs.ctypes.data, n, pad.ctypes.data, value_offsets.ctypes.data,
                                         len(value_offsets) - 1, out.ctypes.data, offs.ctypes.data, C.byref(err))
    if total < 0:
        return int(total), err.value, None, None
    return 0, -1, out[:total].copy(), offs


class KvTopOneConflictIndex:
    """KeyValueStore.typedTopKConflictIndex(k = 1): integer keys instead of strings."""

    def __init__(self, num_leaders):
        self.n = num_leaders
        self.h = lib().fpo_kvci_new(num_leaders)

    def __del__(self):
        try:
            lib().fpo_kvci_free(self.h)
        except Exception:
            pass

    def put(self, command_key, is_set, keys):
        k, n = _iarr(keys)
        lib().fpo_kvci_put(self.h, command_key[0], command_key[1], int(is_set), k, n)

    def put_snapshot(self, command_key):
        lib().fpo_kvci_put_snapshot(self.h, command_key[0], command_key[1])

    def top_one_conflicts(self, is_set, keys):
        k, n = _iarr(keys)
        out = (C.c_int * self.n)()
        lib().fpo_kvci_top_one_conflicts(self.h, int(is_set), k, n, out)
        return list(out)


class TarjanDependencyGraph:
    """depgraph.TarjanDependencyGraph over int keys: commit / updateExecuted / executeByComponent."""

    def __init__(self):
        self.h = lib().fpo_dg_new()

    def __del__(self):
        try:
            lib().fpo_dg_free(self.h)
        except Exception:
            pass

    def commit(self, key, seq, deps):
        d, n = _iarr(deps)
        lib().fpo_dg_commit(self.h, key, seq, d, n)

    def update_executed(self, keys):
        k, n = _iarr(keys)
        lib().fpo_dg_update_executed(self.h, k, n)

    def execute_by_component(self, num_blockers=-1, cap=1 << 16):
        out = (C.c_int * cap)(); sizes = (C.c_int * cap)(); bl = (C.c_int * cap)(); nb = C.c_int(0)
        nc = lib().fpo_dg_execute_by_component(self.h, num_blockers, out, sizes, cap, bl, C.byref(nb))
        comps, k = [], 0
        for c in range(nc):
            comps.append(list(out[k:k + sizes[c]])); k += sizes[c]
        return comps, sorted(bl[:nb.value])
