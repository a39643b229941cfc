"""Python handle over the EPaxos entry points of include/fpx.h (one replica's cmdLog +
leaderStates; shared/src/main/scala/frankenpaxos/epaxos/Replica.scala).  Messages are
int32 matrices, one row per message, dependency sets dense watermark vectors."""
import ctypes as C

import numpy as np

from . import _lib
from .engine import FpxError

REPLY_NONE, REPLY_OK, REPLY_NACK, REPLY_COMMIT = 0, 1, 2, 3
EV_NONE, EV_FAST_COMMIT, EV_SLOW_ACCEPT, EV_TIMER, EV_COMMIT = 0, 1, 2, 3, 4


class EpaxosConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("struct_size", "f", "replica_index", "instances_per_replica",
                                         "max_batch", "device")]


def _bind(L):
    if getattr(L, "_ep_bound", False):
        return
    vp, i32, p = C.c_void_p, C.c_int32, C.POINTER
    L.fpx_epaxos_create.argtypes = [p(vp), p(EpaxosConfig)]; L.fpx_epaxos_create.restype = i32
    L.fpx_epaxos_destroy.argtypes = [vp]; L.fpx_epaxos_destroy.restype = None
    L.fpx_epaxos_lead.argtypes = [vp, vp, i32, p(C.c_int64)]; L.fpx_epaxos_lead.restype = i32
    for nm in ("fpx_epaxos_preaccept", "fpx_epaxos_accept", "fpx_epaxos_preacceptok", "fpx_epaxos_acceptok"):
        f = getattr(L, nm)
        f.argtypes = [vp, vp, i32, vp, p(C.c_int64)]; f.restype = i32
    L.fpx_epaxos_entry.argtypes = [vp, i32, i32, vp, p(i32), vp]; L.fpx_epaxos_entry.restype = i32
    L.fpx_depset_union.argtypes = [i32, vp, vp, vp, i32, vp, i32, vp, vp, vp]; L.fpx_depset_union.restype = i32
    L.fpx_epaxos_last_kernel_ms.argtypes = [vp]; L.fpx_epaxos_last_kernel_ms.restype = C.c_float
    L.fpx_depset_union_dense_dev.argtypes = [i32, vp, i32, i32, i32, vp, vp]; L.fpx_depset_union_dense_dev.restype = i32
    L.fpx_epaxos_stream.argtypes = [vp]; L.fpx_epaxos_stream.restype = vp
    L.fpx_epaxos_sync.argtypes = [vp, p(C.c_int64)]; L.fpx_epaxos_sync.restype = i32
    L.fpx_epaxos_lead_dev.argtypes = [vp, vp, i32]; L.fpx_epaxos_lead_dev.restype = i32
    for nm in ("fpx_epaxos_preaccept_dev", "fpx_epaxos_accept_dev", "fpx_epaxos_preacceptok_dev", "fpx_epaxos_acceptok_dev"):
        f = getattr(L, nm)
        f.argtypes = [vp, vp, i32, vp]; f.restype = i32
    L.fpx_epaxos_preaccept_sets.argtypes = [vp, vp, i32, vp, vp, p(C.c_int64)]; L.fpx_epaxos_preaccept_sets.restype = i32
    L._ep_bound = True


class EpaxosReplica:
    def __init__(self, f, replica_index, instances_per_replica, max_batch=1 << 16, device=0):
        self._L = _lib.lib()
        _bind(self._L)
        cfg = EpaxosConfig(C.sizeof(EpaxosConfig), f, replica_index, instances_per_replica, max_batch, device)
        self.f, self.n, self.index = f, 2 * f + 1, replica_index
        self.h = C.c_void_p()
        st = self._L.fpx_epaxos_create(C.byref(self.h), C.byref(cfg))
        if st != 0:
            self.h = None
            raise FpxError(st)

    def close(self):
        if getattr(self, "h", None):
            self._L.fpx_epaxos_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, fn, rows, width, out_width):
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, width)
        out = np.zeros((max(len(rows), 1), max(out_width, 1)), dtype=np.int32)
        err = C.c_int64(-1)
        if out_width:
            st = fn(self.h, rows.ctypes.data, len(rows), out.ctypes.data, C.byref(err))
        else:
            st = fn(self.h, rows.ctypes.data, len(rows), C.byref(err))
        if st != 0:
            raise FpxError(st, err.value)
        return out[:len(rows)]

    def lead(self, rows):
        """rows: {inst_replica, inst_number, b_ord, b_rep, value_id, seq, avoid_fast_path, 0, deps[n]}"""
        self._call(self._L.fpx_epaxos_lead, rows, 8 + self.n, 0)

    def preaccept(self, rows):
        """rows: {inst_replica, inst_number, b_ord, b_rep, value_id, seq, local_deps[n], msg_deps[n]}
        -> replies {kind, b_ord, b_rep, seq, deps[n]}"""
        return self._call(self._L.fpx_epaxos_preaccept, rows, 6 + 2 * self.n, 4 + self.n)

    def accept(self, rows):
        return self._call(self._L.fpx_epaxos_accept, rows, 6 + self.n, 4 + self.n)

    def preacceptok(self, rows):
        """rows: {inst_replica, inst_number, b_ord, b_rep, from, seq, deps[n]} -> events {kind, seq, deps[n]}"""
        return self._call(self._L.fpx_epaxos_preacceptok, rows, 6 + self.n, 2 + self.n)

    def acceptok(self, rows):
        return self._call(self._L.fpx_epaxos_acceptok, rows, 6, 2 + self.n)

    def last_kernel_ms(self):
        return float(self._L.fpx_epaxos_last_kernel_ms(self.h))

    # -- device-pointer calls (asynchronous on self.stream; errors collected by sync())
    @property
    def stream(self):
        return self._L.fpx_epaxos_stream(self.h)

    def sync(self):
        err = C.c_int64(-1)
        st = self._L.fpx_epaxos_sync(self.h, C.byref(err))
        if st != 0:
            raise FpxError(st, err.value)

    def lead_dev(self, d_in, n):
        st = self._L.fpx_epaxos_lead_dev(self.h, d_in, n)
        if st != 0:
            raise FpxError(st)

    def _dev(self, fn, d_in, n, d_out):
        st = fn(self.h, d_in, n, d_out)
        if st != 0:
            raise FpxError(st)

    def preaccept_dev(self, d_in, n, d_reply):
        self._dev(self._L.fpx_epaxos_preaccept_dev, d_in, n, d_reply)

    def accept_dev(self, d_in, n, d_reply):
        self._dev(self._L.fpx_epaxos_accept_dev, d_in, n, d_reply)

    def preacceptok_dev(self, d_in, n, d_event):
        self._dev(self._L.fpx_epaxos_preacceptok_dev, d_in, n, d_event)

    def acceptok_dev(self, d_in, n, d_event):
        self._dev(self._L.fpx_epaxos_acceptok_dev, d_in, n, d_event)

    def preaccept_sets(self, rows, overflow_count):
        """handlePreAccept for possibly sparse sets: FPX_ERR_UNSUPPORTED at the first message that carries
        overflow values (this handle computes with dense watermark vectors)."""
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 6 + 2 * self.n)
        oc = np.ascontiguousarray(overflow_count, dtype=np.int32)
        out = np.zeros((max(len(rows), 1), 4 + self.n), dtype=np.int32)
        err = C.c_int64(-1)
        st = self._L.fpx_epaxos_preaccept_sets(self.h, rows.ctypes.data, len(rows), oc.ctypes.data, out.ctypes.data, C.byref(err))
        if st != 0:
            raise FpxError(st, err.value)
        return out[:len(rows)]

    def entry(self, rep, num):
        out = np.zeros(7 + self.n, dtype=np.int32)
        lk = C.c_int32(0)
        lb = np.zeros(2, dtype=np.int32)
        st = self._L.fpx_epaxos_entry(self.h, rep, num, out.ctypes.data, C.byref(lk), lb.ctypes.data)
        if st != 0:
            raise FpxError(st)
        return out, lk.value, lb


def depset_union(watermarks, value_lists, group_off, device=0):
    """Batched IntPrefixSet union on the GPU.  watermarks[j], value_lists[j] (iterable of
    ints > watermark); group q unions sets group_off[q]..group_off[q+1].  Returns a list of
    (watermark, sorted values)."""
    L = _lib.lib()
    _bind(L)
    wm = np.ascontiguousarray(watermarks, dtype=np.int32)
    off = np.zeros(len(wm) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(v) for v in value_lists])
    vals = np.ascontiguousarray([x for v in value_lists for x in v], dtype=np.int32) if off[-1] else np.zeros(1, np.int32)
    goff = np.ascontiguousarray(group_off, dtype=np.int32)
    ng = len(goff) - 1
    ow, on = np.zeros(max(ng, 1), np.int32), np.zeros(max(ng, 1), np.int32)
    ov = np.zeros(max(int(off[-1]), 1), np.int32)
    st = L.fpx_depset_union(device, wm.ctypes.data, off.ctypes.data, vals.ctypes.data, len(wm), goff.ctypes.data, ng,
                            ow.ctypes.data, on.ctypes.data, ov.ctypes.data)
    if st != 0:
        raise FpxError(st)
    res = []
    for q in range(ng):
        o = off[goff[q]]
        res.append((int(ow[q]), ov[o:o + on[q]].tolist()))
    return res


def depset_union_dense_dev(d_in, n_groups, sets_per_group, n_replicas, d_out, stream=None, device=0):
    """Dense batched union on device pointers: out[q][k] = max_r in[q][r][k] (asynchronous)."""
    L = _lib.lib()
    _bind(L)
    st = L.fpx_depset_union_dense_dev(device, d_in, n_groups, sets_per_group, n_replicas, d_out, stream)
    if st != 0:
        raise FpxError(st)


# --------------------------------------------------------------------------- execution side
def _bind_f4(L):
    if getattr(L, "_f4_bound", False):
        return
    vp, i32, p = C.c_void_p, C.c_int32, C.POINTER
    L.fpx_conflict_index_create.argtypes = [p(vp), i32, i32, i32, i32, i32]; L.fpx_conflict_index_create.restype = i32
    L.fpx_conflict_index_destroy.argtypes = [vp]; L.fpx_conflict_index_destroy.restype = None
    L.fpx_conflict_index_put_snapshot.argtypes = [vp, i32, i32]; L.fpx_conflict_index_put_snapshot.restype = i32
    L.fpx_conflict_index_batch.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, p(C.c_int64)]
    L.fpx_conflict_index_batch.restype = i32
    L.fpx_depgraph_create.argtypes = [p(vp), i32, i32, i32, i32]; L.fpx_depgraph_create.restype = i32
    L.fpx_depgraph_destroy.argtypes = [vp]; L.fpx_depgraph_destroy.restype = None
    L.fpx_depgraph_commit.argtypes = [vp, vp, vp, vp, vp, i32, p(C.c_int64)]; L.fpx_depgraph_commit.restype = i32
    L.fpx_depgraph_update_executed.argtypes = [vp, vp, i32, p(C.c_int64)]; L.fpx_depgraph_update_executed.restype = i32
    L.fpx_depgraph_execute.argtypes = [vp, vp, vp, p(i32), vp]; L.fpx_depgraph_execute.restype = i32
    L._f4_bound = True


class ConflictIndex:
    """KeyValueStore's top-1 conflict index on the GPU (include/fpx.h, fpx_conflict_index_*)."""
    QUERY_THEN_PUT, PUT, QUERY = 0, 1, 2

    def __init__(self, num_leaders, key_capacity=1 << 16, max_commands=1 << 16, max_keys=1 << 17, device=0):
        self._L = _lib.lib()
        _bind_f4(self._L)
        self.n = num_leaders
        self.h = C.c_void_p()
        st = self._L.fpx_conflict_index_create(C.byref(self.h), num_leaders, key_capacity, max_commands, max_keys, device)
        if st != 0:
            self.h = None
            raise FpxError(st)

    def close(self):
        if getattr(self, "h", None):
            self._L.fpx_conflict_index_destroy(self.h)
            self.h = None

    __del__ = close

    def put_snapshot(self, leader, id_):
        st = self._L.fpx_conflict_index_put_snapshot(self.h, leader, id_)
        if st != 0:
            raise FpxError(st)

    def batch(self, leader, id_, is_set, key_lists, mode=0):
        """Commands in delivery order; key_lists[i] = the keys of command i.  Returns deps [n_cmd, num_leaders]
        (None for mode PUT)."""
        n = len(leader)
        leader = np.ascontiguousarray(leader, dtype=np.int32); id_ = np.ascontiguousarray(id_, dtype=np.int32)
        is_set = np.ascontiguousarray(is_set, dtype=np.uint8)
        off = np.zeros(n + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(k) for k in key_lists])
        keys = np.ascontiguousarray(np.concatenate([np.asarray(k, dtype=np.int32) for k in key_lists])
                                    if off[-1] else np.zeros(0, dtype=np.int32), dtype=np.int32)
        out = np.zeros((max(n, 1), self.n), dtype=np.int32)
        err = C.c_int64(-1)
        st = self._L.fpx_conflict_index_batch(self.h, leader.ctypes.data, id_.ctypes.data, is_set.ctypes.data,
                                              off.ctypes.data, keys.ctypes.data if len(keys) else None, n, mode,
                                              out.ctypes.data, C.byref(err))
        if st != 0:
            raise FpxError(st, err.value)
        return None if mode == self.PUT else out[:n]


class DependencyGraph:
    """TarjanDependencyGraph on the GPU (include/fpx.h, fpx_depgraph_*): int keys in [0, key_capacity)."""

    def __init__(self, key_capacity=1 << 12, dep_pool_capacity=1 << 16, max_batch=1 << 12, device=0):
        self._L = _lib.lib()
        _bind_f4(self._L)
        self.cap = key_capacity
        self.h = C.c_void_p()
        st = self._L.fpx_depgraph_create(C.byref(self.h), key_capacity, dep_pool_capacity, max_batch, device)
        if st != 0:
            self.h = None
            raise FpxError(st)

    def close(self):
        if getattr(self, "h", None):
            self._L.fpx_depgraph_destroy(self.h)
            self.h = None

    __del__ = close

    def commit(self, keys, seqs, dep_lists):
        n = len(keys)
        keys = np.ascontiguousarray(keys, dtype=np.int32); seqs = np.ascontiguousarray(seqs, dtype=np.int32)
        off = np.zeros(n + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(d) for d in dep_lists])
        deps = np.ascontiguousarray(np.concatenate([np.asarray(d, dtype=np.int32) for d in dep_lists])
                                    if off[-1] else np.zeros(0, dtype=np.int32), dtype=np.int32)
        err = C.c_int64(-1)
        st = self._L.fpx_depgraph_commit(self.h, keys.ctypes.data, seqs.ctypes.data, off.ctypes.data,
                                         deps.ctypes.data if len(deps) else None, n, C.byref(err))
        if st != 0:
            raise FpxError(st, err.value)

    def update_executed(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.int32)
        err = C.c_int64(-1)
        st = self._L.fpx_depgraph_update_executed(self.h, keys.ctypes.data, len(keys), C.byref(err))
        if st != 0:
            raise FpxError(st, err.value)

    def execute_by_component(self):
        """(components in dependency order, each a list of keys sorted by (seq, key); sorted blockers)."""
        out = np.zeros(self.cap, dtype=np.int32); head = np.zeros(self.cap, dtype=np.uint8)
        bl = np.zeros(self.cap, dtype=np.uint8)
        n = C.c_int32(0)
        st = self._L.fpx_depgraph_execute(self.h, out.ctypes.data, head.ctypes.data, C.byref(n), bl.ctypes.data)
        if st != 0:
            raise FpxError(st)
        comps = []
        for p in range(n.value):
            if head[p]:
                comps.append([])
            comps[-1].append(int(out[p]))
        return comps, np.nonzero(bl)[0].tolist()
