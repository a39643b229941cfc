"""Slot-residue sharding across the GPUs of one box.

Rank g of P owns the slots with slot % P == g -- the reference's own way of
partitioning a log (`slot % numAcceptorGroups`,
shared/src/main/scala/frankenpaxos/multipaxos/ProxyLeader.scala:190) -- and runs
its own engine on them with no data-path collective.  The one exchange a single
global log needs is the executable prefix: replicas execute in slot order and
stop at the first hole (multipaxos/Replica.scala:397-402), so every rank
contributes its first not-yet-chosen GLOBAL slot and the global watermark is the
minimum.  Pure torch / numpy: runs under gloo on CPU and nccl on GPU.
"""
import numpy as np


def owner(slot, world):
    return np.asarray(slot) % world


def split(records, world, field="slot"):
    """Records of each rank, delivery order preserved within a rank."""
    k = records[field] % world
    return [records[k == g] for g in range(world)]


def local_to_global(local_index, rank, world):
    return np.asarray(local_index, dtype=np.int64) * world + rank


def global_watermark(first_unchosen_global_slot, group=None):
    """all_gather one integer per rank, return (min over ranks, gathered tensor).
    `first_unchosen_global_slot` is a 1-element int32 tensor on the rank's device."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(first_unchosen_global_slot.item()), first_unchosen_global_slot.clone()
    world = dist.get_world_size(group)
    out = torch.empty(world, dtype=first_unchosen_global_slot.dtype, device=first_unchosen_global_slot.device)
    dist.all_gather_into_tensor(out, first_unchosen_global_slot.reshape(1), group=group)
    return int(out.min().item()), out


def connect(engine, group=None):
    """Wire the engines of a sharded log (one per rank) together: every rank exports the IPC handle of
    its frontier table, the handles are all-gathered (plumbing: torch.distributed, any backend), and
    every rank attaches every peer's table.  After this, each engine's watermark publication lands in
    all tables over NVLink from inside the publishing kernel (include/fpx.h, fpx_exchange_*), and
    `engine.global_watermark()` is a local read."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.frombuffer(bytearray(engine.exchange_export()), dtype=torch.uint8).clone()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    gathered = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(gathered, mine.to(dev), group=group)
    handles = [bytes(t.cpu().numpy().tobytes()) for t in gathered]
    for shard, h in enumerate(handles):
        if shard != rank:
            engine.exchange_attach(shard, h)
    dist.barrier(group=group)        # every table is attached everywhere before anyone publishes
    return handles


def min_over_shards(values, group=None):
    """Element-wise minimum of an int32 numpy vector over the ranks (MIN all-reduce; identity on one rank)."""
    import torch
    import torch.distributed as dist
    values = np.ascontiguousarray(values, dtype=np.int32)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return values
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(values.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t.cpu().numpy()


def replica_chosen_range(engine, recs, group=None):
    """Replica.handleChosenNoopRange (mencius/Replica.scala:464-486) on a sharded log: the handler stops at the
    first slot of the range that is already in the log, and that slot may belong to another shard -- every rank
    reports its first hit per record, the minimum over the ranks bounds every rank's fill."""
    first = engine.mencius_replica_range_first(recs)
    engine.mencius_replica_range_fill(recs, min_over_shards(first, group))


def merge_chosen_in_slot_order(per_rank_chosen):
    """A replica's view of the sharded log: Chosen records of all ranks by slot."""
    allc = np.concatenate(per_rank_chosen)
    return allc[np.argsort(allc["slot"], kind="stable")]
