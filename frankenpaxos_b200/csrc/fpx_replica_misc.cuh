#pragma once
#include "fpx_common.cuh"

namespace fpx {

// Sequence numbers are 32-bit; before they wrap, every recorded first-delivery
// stamp is collapsed to 0 ("before everything"): a key that is still pending
// has fewer old stamps than a quorum, so their relative order can never decide
// a completing vote again.
__global__ void renormalize_stamps_kernel(Geometry g, uint32_t* rows, size_t n_rows) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  uint32_t* r = rows + i * g.row_words;
  for (int v = 0; v < g.voters; ++v)
    if (r[2 + v] != kStampEmpty) r[2 + v] = 0;
}

// ===========================================================================
// K5  Replica.handleChosen   S/multipaxos/Replica.scala:572-588
//   log.get(slot): Some -> redundant, ignore; None -> log.put.  First Chosen in
//   delivery order wins: 64-bit atomicMin of (seq : value_id).
// K6  executeLog's prefix rule (:394-402): first hole at or after the watermark.
// ===========================================================================
struct ReplicaParams {
  Geometry g;
  const int2* in;
  int32_t n;               // < 0: read the count from st->n_chosen
  uint32_t seq_base;
  unsigned long long* rlog;
  DevStatus* st;
};

constexpr int kReplicaUnroll = 4;

__global__ void __launch_bounds__(256) replica_chosen_kernel(ReplicaParams P) {
  const Geometry& g = P.g;
  __shared__ int s_mx[8];
  const int n = P.n >= 0 ? P.n : __ldcg(&P.st->n_chosen);
  const int stride = gridDim.x * blockDim.x;
  int mx = INT_MIN;
  // kReplicaUnroll independent records (and their atomics) in flight per thread
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * kReplicaUnroll) {
    int2 rec[kReplicaUnroll];
#pragma unroll
    for (int u = 0; u < kReplicaUnroll; ++u) {
      int i = i0 + u * stride;
      rec[u] = i < n ? __ldcg(P.in + i) : make_int2(-1, 0);
    }
#pragma unroll
    for (int u = 0; u < kReplicaUnroll; ++u) {
      int i = i0 + u * stride;
      if (i >= n) break;
      int local = local_slot(g, rec[u].x);
      if (local < 0) {   // a retired slot is in the log (and executed) already: redundantly chosen (:580-586)
        if (local != kLocalRetired) report_error(P.st, FPX_ERR_SLOT_RANGE, i);
        continue;
      }
      // first Chosen per slot wins (Replica.scala:580-588): min over (delivery seq : value)
      red_min_u64(&P.rlog[local], ((unsigned long long)(P.seq_base + (uint32_t)i) << 32) | (uint32_t)rec[u].y);
      mx = max(mx, g.base_local + ring_to_rel(g, local));   // ordinal
    }
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = max(mx, s_mx[w]);
    if (mx != INT_MIN) atomicMax(&P.st->max_chosen_local, mx);   // one same-address atomic per CTA
  }
}

// executeLog's prefix rule (Replica.scala:394-402): the first hole at or after the
// watermark.  Coalesced scan of the replica log over [watermark, last chosen + 2); the
// last CTA to finish publishes the result.
__global__ void __launch_bounds__(256) watermark_scan_kernel(Geometry g, const unsigned long long* rlog,
                                                            DevStatus* st, int32_t* d_out, DevExchange* xch) {
  __shared__ int s_found[8];
  __shared__ bool s_last;
  const int lo = __ldcg(&st->wm_local);                                                    // ordinals
  const int hi = min(__ldcg(&st->max_chosen_local) + 2, g.base_local + g.local_slots);      // one past the last candidate hole
  int found = INT_MAX;
  // four independent loads per round trip (the early exit makes consecutive iterations dependent)
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi && found == INT_MAX; i += 4 * stride) {
    unsigned long long v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i + u * stride < hi) ? __ldcg(&rlog[rel_to_ring(g, i + u * stride - g.base_local)]) : 0ull;
#pragma unroll
    for (int u = 3; u >= 0; --u)
      if (i + u * stride < hi && v[u] == kU64Empty) found = (int)(i + u * stride);   // ascending: the smallest one last
  }
  found = __reduce_min_sync(0xffffffffu, found);
  if ((threadIdx.x & 31) == 0) s_found[threadIdx.x >> 5] = found;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) found = min(found, s_found[w]);
    if (found != INT_MAX) atomicMin(&st->wm_found, found);
    __threadfence();
    s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  int f = min(*(volatile int*)&st->wm_found, hi);
  f = max(f, lo);
  if (f > g.base_local + g.local_slots) f = g.base_local + g.local_slots;
  st->wm_local = f;
  st->wm_found = INT_MAX;
  st->ticket = 0;
  int global = f * g.shard_count + g.shard_index;
  st->watermark = global;
  if (d_out) *d_out = global;
  exchange_publish(xch, global);
}

// Global executable prefix of a sharded log as of publication `epoch` of every shard: lane p waits until
// entry p of this engine's table carries an epoch >= `epoch` (peers store into it over NVLink), then the
// minimum (Replica.scala:397-402: execution stops at the first hole of the GLOBAL log).  Bounded wait.
__global__ void global_watermark_kernel(DevExchange* x, const unsigned long long* table, uint32_t epoch,
                                        int32_t* d_out, int32_t* d_frontiers, DevStatus* st,
                                        unsigned long long timeout_ns) {
  const int lane = threadIdx.x;
  int mine = INT_MAX;
  bool late = false;
  const unsigned long long t0 = global_timer_ns();
  for (int p = lane; p < x->n; p += 32) {
    unsigned long long v;
    do {
      asm volatile("ld.global.acquire.sys.u64 %0, [%1];" : "=l"(v) : "l"(table + p) : "memory");
      if ((int32_t)((uint32_t)(v >> 32) - epoch) >= 0) break;
      if (global_timer_ns() - t0 > timeout_ns) { late = true; break; }
      __nanosleep(200);
    } while (true);
    if (d_frontiers) d_frontiers[p] = (int32_t)(uint32_t)v;
    mine = min(mine, (int32_t)(uint32_t)v);
  }
  mine = __reduce_min_sync(0xffffffffu, mine);
  if (__any_sync(0xffffffffu, late) && lane == 0) report_error(st, FPX_ERR_EXCHANGE_TIMEOUT, 0);
  if (lane == 0 && d_out) *d_out = mine;
}
__global__ void renormalize_rlog_kernel(unsigned long long* rlog, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && rlog[i] != kU64Empty) rlog[i] &= 0xffffffffull;
}

// ===========================================================================
// K7  batched quorum predicates  S/quorums/Grid.scala:35-56,
//     S/quorums/SimpleMajority.scala:41-55
// ===========================================================================
__global__ void quorum_eval_kernel(Geometry g, int which, const uint32_t* masks, int n, uint8_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t m = masks[i];
  bool foreign = (m >> 31) & 1u;
  int members = g.flexible ? g.groups * g.per_group : g.per_group;
  uint32_t member_mask = members >= 31 ? 0x7fffffffu : ((1u << members) - 1u);
  if (m & ~member_mask & 0x7fffffffu) foreign = true;
  m &= member_mask;
  uint8_t r;
  if (g.flexible) {
    bool rd = read_quorum(g, m), wr = write_quorum(g, m);
    r = (which == 0 || which == 2) ? rd : wr;
  } else {
    int q = members / 2 + 1;  // SimpleMajority.scala:30
    r = __popc(m) >= q;
  }
  if (foreign && which < 2) r = 2;  // `require(nodes.subsetOf(members))` throws
  out[i] = r;
}

// ---- snapshots (Phase1b / parity read-back; not on the hot path)
__global__ void snapshot_votes_kernel(Geometry g, const unsigned long long* votes, int group, int acceptor,
                                      int first_slot, int n_slots, int32_t* vote_round, int32_t* vote_value) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  int slot = first_slot + i;
  int l = local_slot(g, slot);
  int v = l >= 0 ? voter_index(g, group, acceptor, slot) : -1;
  if (!g.flexible && v >= 0 && group != expected_group(g, slot)) v = -1;  // this acceptor never sees that slot
  int vr = -1, vv = -1;
  if (v >= 0) {
    unsigned long long c = votes[cell_index(g, l, v)];
    if ((c >> 32) != 0) { vr = (int)(c >> 32) - 1; vv = (int)(uint32_t)c; }
  }
  vote_round[i] = vr;
  vote_value[i] = vv;
}
// fpx_retire_below: the ring positions of the ordinals [first, first + count) go back to their initial state
// (row unarmed, no votes, no log entry) so that they can serve the slots one window ahead.
__global__ void recycle_kernel(Geometry g, uint32_t* rows, unsigned long long* votes, unsigned long long* rlog,
                               int first_rel, int count) {
  const int per_row = g.row_words + 2 * (g.voters << g.cell_shift) + 2;   // 32-bit words of state per slot
  const long long total = (long long)count * per_row;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t / per_row), w = (int)(t % per_row);
    const int ring = rel_to_ring(g, first_rel + k);
    if (w < g.row_words) {
      rows[(size_t)ring * g.row_words + w] = 0xffffffffu;
    } else if (w < g.row_words + 2 * (g.voters << g.cell_shift)) {
      const int c = w - g.row_words;                       // 32-bit half of a 64-bit cell / claim word
      const bool claim = g.cell_shift && ((c >> 1) & 1);   // vanilla: odd 64-bit words are the batch claims (~0)
      ((uint32_t*)(votes + ((size_t)ring * g.voters << g.cell_shift)))[c] = claim ? 0xffffffffu : 0u;
    } else {
      ((uint32_t*)(rlog + ring))[w - g.row_words - 2 * (g.voters << g.cell_shift)] = 0xffffffffu;
    }
  }
}

// Leader.safeValue over a Phase-1 quorum (S/multipaxos/Leader.scala:318-329): coalesced
// sweep of the flat slot x voter cell array, arg-max of voteRound per slot over the
// responders' cells (64-bit max of {voteRound+1, value}); warp redux for maxPhase1bSlot.
__global__ void __launch_bounds__(256) safe_values_kernel(Geometry g, const unsigned long long* votes,
                                                         uint32_t responders, int first_slot, int n_slots,
                                                         int32_t* vote_round, int32_t* value, int32_t* d_max_slot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int best_slot = -1;
  if (i < n_slots) {
    int slot = first_slot + i;
    int l = local_slot(g, slot);
    unsigned long long m = 0;
    if (l >= 0) {
      int grp = expected_group(g, slot);
      for (int v = 0; v < g.voters; ++v) {
        int gid = g.flexible ? v : grp * g.per_group + v;   // global acceptor id of voter v of this slot
        if (!((responders >> gid) & 1u)) continue;
        unsigned long long c = __ldcg(&votes[cell_index(g, l, v)]);
        if (c & kCellChosen) continue;
        m = max(m, c);
      }
    }
    bool any = (m >> 32) != 0;
    vote_round[i] = any ? (int)(m >> 32) - 1 : -1;
    value[i] = any ? (int)(uint32_t)m : -1;                // else Noop (:325)
    if (any) best_slot = slot;
  }
  best_slot = __reduce_max_sync(0xffffffffu, best_slot);
  if ((threadIdx.x & 31) == 0 && best_slot >= 0) atomicMax(d_max_slot, best_slot);
}

__global__ void snapshot_log_kernel(Geometry g, const unsigned long long* rlog, int first_slot, int n_slots,
                                    int32_t* value) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  int l = local_slot(g, first_slot + i);
  int v = -1;
  if (l >= 0 && rlog[l] != kU64Empty) v = (int)(uint32_t)rlog[l];
  value[i] = v;
}

}  // namespace fpx
