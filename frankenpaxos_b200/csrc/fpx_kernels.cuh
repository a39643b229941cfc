// fpx_kernels.cuh -- sm_100a kernels of the quorum-vote engine.
//
// Everything here is integer scatter/gather + order-preserving scans bounded by
// HBM/L2 bandwidth; there is no GEMM-shaped work, hence no tcgen05/TMEM.  What
// matters (DESIGN.md "kernels"): one 128-bit load per message record, fully
// coalesced striped tiles, at most one 32-byte sector touched per random state
// access (a proxy-leader row IS one sector for <= 6 voters), warp-level
// ballots/shuffles for the in-order prefix logic, single-pass decoupled
// look-back for every cross-tile dependency (no second read of the stream).
//
// Reference semantics each kernel reproduces are cited at the kernel.
// S/ = shared/src/main/scala/frankenpaxos/ in mwhittaker/frankenpaxos.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fpx.h"

namespace fpx {

// ---------------------------------------------------------------------------
// constants / layouts
// ---------------------------------------------------------------------------
constexpr uint32_t kUnarmed = 0xffffffffu;      // row.round_word of a never-armed key
constexpr uint32_t kDoneBit = 0x80000000u;      // row.round_word bit: ProxyLeader `Done`
constexpr uint32_t kStampEmpty = 0xffffffffu;   // no Phase2b from this voter yet
constexpr uint64_t kU64Empty = ~0ull;
constexpr int kTileThreads = 256;
constexpr int kTileItems = 4;
constexpr int kTile = kTileThreads * kTileItems;  // records per tile
constexpr int kChunks = kTile / 32;               // 32-record warp chunks per tile
constexpr int kMaxKeys = FPX_MAX_ACCEPTORS;       // acceptors tracked by the round scan
constexpr int kMaxConflicts = 1024;

// Device-resident status block (one per engine).
struct DevStatus {
  unsigned long long err_word;  // min over errors of (index << 8 | -code); ~0 = none
  int32_t n_p2b, n_nack, n_chosen, watermark;
  uint32_t n_conflicts;         // entries in the conflict list of the running call
  uint32_t ticket;              // last-block-done ticket
  int32_t wm_local;             // replica: first local index not yet chosen
  int32_t max_chosen_local;     // replica: largest local index ever chosen
  int32_t wm_found;             // scratch of the watermark scan
  int32_t pad;
};

struct Geometry {
  int32_t protocol, f, groups, per_group, flexible, num_leaders;
  int32_t voters;          // acceptors that can vote on one slot (row width)
  int32_t num_keys;        // total acceptors = groups * per_group
  int32_t quorum;          // f + 1 (count predicate)
  int32_t row_words;       // 8 / 16 / 32 uint32 per proxy-leader row
  int32_t slot_capacity;   // global
  int32_t local_slots;     // rows held by this shard
  int32_t shard_index, shard_count;
  uint32_t ovf_mask;       // overflow_capacity - 1, or 0 with ovf_cap == 0
  int32_t ovf_cap;
};

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  // streaming 128-bit load: message records are read exactly once
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(int4* p, int4 v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ void st_stream2(int2* p, int2 v) {
  asm volatile("st.global.L1::no_allocate.v2.s32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y));
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ void report_error(DevStatus* st, int code, long long index) {
  unsigned long long w = ((unsigned long long)index << 8) | (unsigned long long)(unsigned)(-code);
  atomicMin(&st->err_word, w);
}

// global slot -> local row index of this shard, or -1
__device__ __forceinline__ int local_slot(const Geometry& g, int slot) {
  if (slot < 0 || slot >= g.slot_capacity) return -1;
  if (g.shard_count == 1) return slot;
  if (slot % g.shard_count != g.shard_index) return -1;
  return slot / g.shard_count;
}

// (group, acceptor) -> voter index within the slot's row, or -1.
// non-flexible: the slot's group is slot % numAcceptorGroups
// (S/multipaxos/ProxyLeader.scala:190); flexible: every (row, col) of the grid
// (S/multipaxos/ProxyLeader.scala:118-124).
__device__ __forceinline__ int voter_index(const Geometry& g, int group, int acceptor, int slot) {
  if (group < 0 || group >= g.groups || acceptor < 0 || acceptor >= g.per_group) return -1;
  if (g.flexible) return group * g.per_group + acceptor;
  if (g.protocol == FPX_MULTIPAXOS && group != slot % g.groups) return -1;
  return acceptor;
}

// Quorum predicate over a voter bitmask: non-flexible `size >= f+1`
// (S/multipaxos/ProxyLeader.scala:238-240); flexible Grid.isWriteQuorum, one
// member of every row (S/quorums/Grid.scala:49).
__device__ __forceinline__ bool write_quorum(const Geometry& g, uint32_t mask) {
  if (!g.flexible) return __popc(mask) >= g.quorum;
  uint32_t rowmask = (1u << g.per_group) - 1u;
  for (int r = 0; r < g.groups; ++r) {
    if (((mask >> (r * g.per_group)) & rowmask) == 0) return false;
  }
  return true;
}
__device__ __forceinline__ bool read_quorum(const Geometry& g, uint32_t mask) {
  // Grid.isReadQuorum: some row entirely inside (S/quorums/Grid.scala:40)
  uint32_t rowmask = (1u << g.per_group) - 1u;
  for (int r = 0; r < g.groups; ++r) {
    if (((mask >> (r * g.per_group)) & rowmask) == rowmask) return true;
  }
  return false;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// Proxy-leader row: {u32 round_word; i32 value_id; u32 stamp[voters]; pad}.
// round_word = kUnarmed | round | kDoneBit.  stamp[v] = global sequence number
// of the FIRST Phase2b delivery from voter v for this (slot, round).
struct RowRef {
  uint32_t* p;
  __device__ __forceinline__ uint32_t round_word() const { return p[0]; }
  __device__ __forceinline__ int value_id() const { return (int)p[1]; }
  __device__ __forceinline__ uint32_t* stamps() const { return p + 2; }
  __device__ __forceinline__ unsigned long long* hdr64() const { return (unsigned long long*)p; }
};

struct PLState {
  uint32_t* rows;               // local_slots * row_words
  unsigned long long* ovf_keys; // ovf_cap
  uint32_t* ovf_rows;           // ovf_cap * row_words
};

// Find the row of key (slot, round): the primary row when its armed round
// matches, else the overflow table.  Returns p == nullptr when the key was never
// armed (ProxyLeader.scala:220-225 `case None`).
__device__ __forceinline__ RowRef find_row(const Geometry& g, const PLState& s, int local, int slot, int round) {
  RowRef r{s.rows + (size_t)local * g.row_words};
  uint32_t rw = r.round_word();
  if (rw != kUnarmed && (int)(rw & ~kDoneBit) == round) return r;
  if (rw == kUnarmed || g.ovf_cap == 0) return RowRef{nullptr};
  unsigned long long key = ((unsigned long long)(uint32_t)slot << 32) | (uint32_t)round;
  uint32_t h = (uint32_t)mix64(key) & g.ovf_mask;
  for (int probe = 0; probe < g.ovf_cap; ++probe) {
    unsigned long long k = s.ovf_keys[h];
    if (k == key) return RowRef{s.ovf_rows + (size_t)h * g.row_words};
    if (k == kU64Empty) break;
    h = (h + 1) & g.ovf_mask;
  }
  return RowRef{nullptr};
}

// ---------------------------------------------------------------------------
// Single-value decoupled look-back (exclusive prefix SUM over tiles).
// desc[tile] = (epoch << 34) | (state << 32) | value, state 1 = aggregate,
// 2 = inclusive prefix.  Called by one thread of the tile.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long desc_pack(uint32_t epoch, uint32_t state, uint32_t v) {
  return ((unsigned long long)((epoch << 2) | state) << 32) | v;
}
__device__ __forceinline__ uint32_t lookback_sum(unsigned long long* desc, int tile, uint32_t epoch,
                                                 uint32_t aggregate) {
  if (tile == 0) {
    st_volatile_u64(&desc[0], desc_pack(epoch, 2, aggregate));
    return 0;
  }
  st_volatile_u64(&desc[tile], desc_pack(epoch, 1, aggregate));
  uint32_t excl = 0;
  int t = tile - 1;
  while (true) {
    unsigned long long d = ld_volatile_u64(&desc[t]);
    uint32_t flag = (uint32_t)(d >> 32);
    if ((flag >> 2) != epoch || (flag & 3) == 0) continue;  // not published yet
    excl += (uint32_t)d;
    if ((flag & 3) == 2) break;
    --t;
  }
  st_volatile_u64(&desc[tile], desc_pack(epoch, 2, excl + aggregate));
  return excl;
}
// Keyed MAX look-back: lane k of the calling warp owns key k; desc is
// [tile][kMaxKeys].  Returns the exclusive prefix max (int32) for key k.
__device__ __forceinline__ int lookback_max(unsigned long long* desc, int tile, uint32_t epoch, int key,
                                            int aggregate, int base) {
  unsigned long long* mine = desc + (size_t)tile * kMaxKeys + key;
  if (tile == 0) {
    st_volatile_u64(mine, desc_pack(epoch, 2, (uint32_t)max(base, aggregate)));
    return base;
  }
  st_volatile_u64(mine, desc_pack(epoch, 1, (uint32_t)aggregate));
  int excl = INT_MIN;
  int t = tile - 1;
  while (true) {
    unsigned long long d = ld_volatile_u64(desc + (size_t)t * kMaxKeys + key);
    uint32_t flag = (uint32_t)(d >> 32);
    if ((flag >> 2) != epoch || (flag & 3) == 0) continue;
    excl = max(excl, (int)(uint32_t)d);
    if ((flag & 3) == 2) break;
    --t;
  }
  st_volatile_u64(mine, desc_pack(epoch, 2, (uint32_t)max(excl, aggregate)));
  return excl;
}

// ===========================================================================
// K1  ProxyLeader.handlePhase2a  -- "arm"   S/multipaxos/ProxyLeader.scala:175-215
//   states.get((slot, round)): Some -> ignore (:177-183); None -> Pending(phase2a,
//   {}) (:213).  One thread per record; the key's header {round_word, value_id}
//   is claimed with ONE 64-bit CAS.  A second round for a slot whose primary row
//   is taken goes to the overflow table.  Two arms of one key with DIFFERENT
//   values inside one batch (never produced by a correct leader) are resolved to
//   "first in delivery order wins" by the last block (resolve_arm_conflicts).
// ===========================================================================
struct ArmConflict { int32_t slot, round; };

struct ArmParams {
  Geometry g;
  PLState pl;
  const int4* in;
  int32_t n;
  DevStatus* st;
  ArmConflict* conflicts;
  uint32_t* win_bits;  // ceil(n/32): record i created its key's entry
};

__device__ __forceinline__ void note_arm_conflict(const ArmParams& P, int slot, int round) {
  uint32_t c = atomicAdd(&P.st->n_conflicts, 1u);
  if (c < (uint32_t)kMaxConflicts) P.conflicts[c] = ArmConflict{slot, round};
}

// returns true if this record installed the header
__device__ __forceinline__ bool claim_header(const ArmParams& P, RowRef r, int slot, int round, int value,
                                             bool* other_round) {
  unsigned long long want = ((unsigned long long)(uint32_t)value << 32) | (uint32_t)round;
  unsigned long long old = atomicCAS(r.hdr64(), kU64Empty, want);
  *other_round = false;
  if (old == kU64Empty) return true;
  uint32_t orw = (uint32_t)old;
  if ((int)(orw & ~kDoneBit) != round) { *other_round = true; return false; }
  if ((uint32_t)(old >> 32) != (uint32_t)value) note_arm_conflict(P, slot, round);
  return false;
}

__global__ void __launch_bounds__(256) arm_kernel(ArmParams P) {
  const Geometry& g = P.g;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool won = false;  // this record created the key's Pending entry (:213)
  if (i < P.n) {
    int4 rec = ld_stream(P.in + i);  // {slot, round, value_id, dst}
    int slot = rec.x, round = rec.y, value = rec.z;
    int local = local_slot(g, slot);
    if (local < 0) {
      report_error(P.st, FPX_ERR_SLOT_RANGE, i);
    } else if (round < 0 || round > FPX_MAX_ROUND) {
      report_error(P.st, FPX_ERR_ROUND_RANGE, i);
    } else {
      RowRef r{P.pl.rows + (size_t)local * g.row_words};
      bool other;
      won = claim_header(P, r, slot, round, value, &other);
      if (other) {
        // secondary round of this slot -> overflow table (SURVEY 8(g) rule 3)
        if (g.ovf_cap == 0) {
          report_error(P.st, FPX_ERR_OVERFLOW_FULL, i);
        } else {
          unsigned long long key = ((unsigned long long)(uint32_t)slot << 32) | (uint32_t)round;
          uint32_t h = (uint32_t)mix64(key) & g.ovf_mask;
          bool placed = false;
          for (int probe = 0; probe < g.ovf_cap; ++probe) {
            unsigned long long k = atomicCAS(&P.pl.ovf_keys[h], kU64Empty, key);
            if (k == kU64Empty || k == key) {
              RowRef o{P.pl.ovf_rows + (size_t)h * g.row_words};
              bool dummy;
              won = claim_header(P, o, slot, round, value, &dummy);
              placed = true;
              break;
            }
            h = (h + 1) & g.ovf_mask;
          }
          if (!placed) report_error(P.st, FPX_ERR_OVERFLOW_FULL, i);
        }
      }
    }
  }
  unsigned wb = __ballot_sync(0xffffffffu, won);
  if ((threadIdx.x & 31) == 0 && i < P.n) P.win_bits[i >> 5] = wb;

  // ---- last block: two arms of one key with different values.  If the key was
  // created by a record of THIS batch, the lowest-index arm is the one the
  // reference would have kept (later ones hit `case Some(_)`, :177-183); if it
  // existed before the batch, the stored value stands.
  __shared__ bool s_last;
  __shared__ int s_min, s_any;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&P.st->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  uint32_t nc = *(volatile uint32_t*)&P.st->n_conflicts;
  if (threadIdx.x == 0) P.st->ticket = 0;
  if (nc == 0) return;
  if (nc > (uint32_t)kMaxConflicts) {
    if (threadIdx.x == 0) { report_error(P.st, FPX_ERR_CONFLICT, 0); P.st->n_conflicts = 0; }
    return;
  }
  for (uint32_t c = 0; c < nc; ++c) {
    int slot = P.conflicts[c].slot, round = P.conflicts[c].round;
    if (threadIdx.x == 0) { s_min = INT_MAX; s_any = 0; }
    __syncthreads();
    for (int j = threadIdx.x; j < P.n; j += blockDim.x) {
      int4 rec = P.in[j];
      if (rec.x == slot && rec.y == round) {
        atomicMin(&s_min, j);
        if ((__ldcg(&P.win_bits[j >> 5]) >> (j & 31)) & 1u) s_any = 1;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_any && s_min != INT_MAX) {
      RowRef r = find_row(g, P.pl, local_slot(g, slot), slot, round);
      if (r.p != nullptr) r.p[1] = (uint32_t)P.in[s_min].z;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) P.st->n_conflicts = 0;
}

// ===========================================================================
// K2  Acceptor.handlePhase2a   S/multipaxos/Acceptor.scala:184-220
//   One launch handles the interleaved delivery stream of all acceptors.
//   `round` is ONE scalar per acceptor (:95), so for record i addressed to
//   acceptor k the handler's test (:192) is
//        msg.round < max(round_k at batch start, max_{j<i, dst_j==k} msg_j.round)
//   (rejected messages are below the running max, so including them is
//   harmless): an exclusive keyed prefix-max in delivery order.  In-tile: warp
//   chunks of 32 consecutive records, shuffle scans per distinct key (skipped
//   when the whole chunk carries one round -- the steady state), chunk
//   aggregates in shared memory; across tiles: keyed decoupled look-back, lane
//   k of warp 0 owns acceptor k.  Accepted records then get their position in
//   the Phase2b stream from a second (sum) look-back, so both reply streams
//   come out compacted in delivery order in the same pass.
//   Vote cell: states(slot) = State(round, value) (:205-208) is one 64-bit
//   atomicMax of (round+1 : value_id): accepted rounds never decrease in
//   delivery order, so max == last writer, except "same round, different value"
//   which is flagged and resolved to last-in-order by the last block.
// ===========================================================================
struct VoteConflict { int32_t dst, slot; };

struct AcceptorParams {
  Geometry g;
  const int4* in;
  int32_t n;
  int4* out_p2b;
  int2* out_nack;
  unsigned long long* votes;       // local_slots * voters cells
  int32_t* acc_round;              // num_keys
  int32_t* acc_max_voted;          // num_keys
  uint32_t* accept_bits;           // ceil(n/32)
  unsigned long long* desc_max;    // tiles * kMaxKeys
  unsigned long long* desc_cnt;    // tiles
  uint32_t epoch;
  DevStatus* st;
  VoteConflict* conflicts;
};

__device__ __forceinline__ int warp_incl_scan_max(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v = max(v, o);
  }
  return v;
}

__global__ void __launch_bounds__(kTileThreads) acceptor_phase2a_kernel(AcceptorParams P) {
  const Geometry& g = P.g;
  __shared__ int s_chunk[kChunks][kMaxKeys];  // per chunk per key: aggregate, then exclusive prefix
  __shared__ int s_carry[kMaxKeys];
  __shared__ int s_maxslot[kMaxKeys];
  __shared__ uint32_t s_cnt[kChunks];
  __shared__ uint32_t s_tile_excl;
  __shared__ bool s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int num_tiles = gridDim.x;
  const long long tile_base = (long long)tile * kTile;

  for (int j = tid; j < kChunks * kMaxKeys; j += kTileThreads) (&s_chunk[0][0])[j] = INT_MIN;
  if (tid < kMaxKeys) s_maxslot[tid] = INT_MIN;
  __syncthreads();

  int4 rec[kTileItems];
  int key[kTileItems];   // global acceptor id, -1 = no record / invalid
  int pin[kTileItems];   // exclusive prefix max within the chunk
  int loc[kTileItems];   // local slot
  int vix[kTileItems];   // voter index in the slot's cell row

#pragma unroll
  for (int it = 0; it < kTileItems; ++it) {
    long long i = tile_base + it * kTileThreads + tid;
    key[it] = -1; pin[it] = INT_MIN; loc[it] = -1; vix[it] = -1;
    if (i < P.n) {
      rec[it] = ld_stream(P.in + i);
      int grp = rec[it].w >> 16, acc = rec[it].w & 0xffff;
      int slot = rec[it].x, round = rec[it].y;
      if (grp < 0 || grp >= g.groups || acc >= g.per_group) {
        report_error(P.st, FPX_ERR_BAD_ACCEPTOR, i);
      } else if (round < 0 || round > FPX_MAX_ROUND) {
        report_error(P.st, FPX_ERR_ROUND_RANGE, i);
      } else {
        int l = local_slot(g, slot);
        int v = voter_index(g, grp, acc, slot);
        if (l < 0) report_error(P.st, FPX_ERR_SLOT_RANGE, i);
        else if (v < 0) report_error(P.st, FPX_ERR_BAD_ACCEPTOR, i);
        else { key[it] = grp * g.per_group + acc; loc[it] = l; vix[it] = v; }
      }
    } else {
      rec[it] = make_int4(0, 0, 0, 0);
    }
  }

  // ---- in-chunk keyed exclusive prefix max + chunk aggregates
#pragma unroll
  for (int it = 0; it < kTileItems; ++it) {
    const int c = it * (kTileThreads / 32) + warp;
    const int k = key[it];
    const int r = rec[it].y;
    unsigned valid = __ballot_sync(0xffffffffu, k >= 0);
    if (valid == 0) continue;
    int r0 = __shfl_sync(0xffffffffu, r, __ffs(valid) - 1);
    bool uniform = __all_sync(0xffffffffu, k < 0 || r == r0);
    if (uniform) {
      if (k >= 0) s_chunk[c][k] = r0;  // same value from every writer
    } else {
      unsigned remaining = valid;
      while (remaining) {
        int leader = __ffs(remaining) - 1;
        int kk = __shfl_sync(0xffffffffu, k, leader);
        unsigned m = __ballot_sync(0xffffffffu, k == kk);
        int v = (k == kk) ? r : INT_MIN;
        int incl = warp_incl_scan_max(v, lane);
        int ex = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) ex = INT_MIN;
        if (k == kk) pin[it] = ex;
        if (lane == 31) s_chunk[c][kk] = incl;
        remaining &= ~m;
      }
    }
  }
  __syncthreads();

  // ---- tile scan over chunks + keyed look-back (warp 0, lane = acceptor)
  int incl_k = INT_MIN;
  if (warp == 0) {
    int carry = INT_MIN;
    if (lane < g.num_keys) {
      int run = INT_MIN;
#pragma unroll 4
      for (int c = 0; c < kChunks; ++c) {
        int a = s_chunk[c][lane];
        s_chunk[c][lane] = run;
        run = max(run, a);
      }
      int base = P.acc_round[lane];
      carry = lookback_max(P.desc_max, tile, P.epoch, lane, run, base);
      incl_k = max(carry, run);
      s_carry[lane] = carry;
    }
  }
  __syncthreads();

  // ---- accept decision (Acceptor.scala:192) + position in the reply streams
  bool accept[kTileItems];
  int cur[kTileItems];
  uint32_t rank[kTileItems];
#pragma unroll
  for (int it = 0; it < kTileItems; ++it) {
    const int c = it * (kTileThreads / 32) + warp;
    const int k = key[it];
    cur[it] = INT_MIN;
    accept[it] = false;
    if (k >= 0) {
      cur[it] = max(max(s_carry[k], s_chunk[c][k]), pin[it]);
      accept[it] = rec[it].y >= cur[it];
    }
    unsigned b = __ballot_sync(0xffffffffu, accept[it]);
    rank[it] = __popc(b & lanemask_lt());
    if (lane == 0) {
      s_cnt[c] = __popc(b);
      long long i0 = tile_base + it * kTileThreads + warp * 32;
      if (i0 < P.n) P.accept_bits[i0 >> 5] = b;
    }
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t cnt = s_cnt[lane];  // kChunks == 32
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    s_cnt[lane] = incl - cnt;
    uint32_t tile_total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) {
      uint32_t excl = lookback_sum(P.desc_cnt, tile, P.epoch, tile_total);
      s_tile_excl = excl;
      if (tile == num_tiles - 1) {
        P.st->n_p2b = (int)(excl + tile_total);
        P.st->n_nack = P.n - (int)(excl + tile_total);
      }
    }
    // acceptor round after the batch = inclusive prefix of the last tile (:204)
    if (tile == num_tiles - 1 && lane < g.num_keys) P.acc_round[lane] = incl_k;
  }
  __syncthreads();

  // ---- effects
  const uint32_t tile_excl = s_tile_excl;
#pragma unroll
  for (int it = 0; it < kTileItems; ++it) {
    const int c = it * (kTileThreads / 32) + warp;
    const long long i = tile_base + it * kTileThreads + tid;
    const int k = key[it];
    if (k >= 0) {
      uint32_t pos = tile_excl + s_cnt[c] + rank[it];
      if (accept[it]) {
        int grp = rec[it].w >> 16, acc = rec[it].w & 0xffff;
        // Phase2b(groupIndex, acceptorIndex, slot, round) (:211-219)
        st_stream(P.out_p2b + pos, make_int4(grp, acc, rec[it].x, rec[it].y));
        // states(slot) = State(voteRound = round, voteValue) (:205-208)
        unsigned long long cell = ((unsigned long long)(uint32_t)(rec[it].y + 1) << 32) | (uint32_t)rec[it].z;
        unsigned long long old = atomicMax(&P.votes[(size_t)loc[it] * g.voters + vix[it]], cell);
        if ((old >> 32) == (cell >> 32) && old != cell) {
          uint32_t cidx = atomicAdd(&P.st->n_conflicts, 1u);
          if (cidx < (uint32_t)kMaxConflicts) P.conflicts[cidx] = VoteConflict{rec[it].w, rec[it].x};
        }
      } else {
        // Nack(round = round) to leaders(roundSystem.leader(phase2a.round)) (:197-198)
        uint32_t npos = (uint32_t)i - pos;
        st_stream2(P.out_nack + npos, make_int2(rec[it].y % g.num_leaders, cur[it]));
      }
    }
    // maxVotedSlot = max(maxVotedSlot, slot) (:209): per-key warp reduce
    unsigned remaining = __ballot_sync(0xffffffffu, k >= 0 && accept[it]);
    while (remaining) {
      int leader = __ffs(remaining) - 1;
      int kk = __shfl_sync(0xffffffffu, k, leader);
      unsigned m = __ballot_sync(0xffffffffu, k == kk && accept[it]);
      int v = (k == kk && accept[it]) ? rec[it].x : INT_MIN;
      int mx = __reduce_max_sync(0xffffffffu, v);
      if (lane == leader) atomicMax(&s_maxslot[kk], mx);
      remaining &= ~m;
    }
  }
  __syncthreads();
  if (tid < g.num_keys && s_maxslot[tid] != INT_MIN) atomicMax(&P.acc_max_voted[tid], s_maxslot[tid]);

  // ---- last block: same (acceptor, slot, round) voted twice with different
  // values in one batch -> the later delivery must win (map overwrite, :205)
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&P.st->ticket, 1u) == (uint32_t)num_tiles - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  uint32_t nc = *(volatile uint32_t*)&P.st->n_conflicts;
  if (tid == 0) P.st->ticket = 0;
  if (nc == 0) return;
  if (nc > (uint32_t)kMaxConflicts) {
    if (tid == 0) { report_error(P.st, FPX_ERR_CONFLICT, 0); P.st->n_conflicts = 0; }
    return;
  }
  __shared__ int s_win;
  for (uint32_t cix = 0; cix < nc; ++cix) {
    int dst = P.conflicts[cix].dst, slot = P.conflicts[cix].slot;
    if (tid == 0) s_win = -1;
    __syncthreads();
    for (int j = tid; j < P.n; j += kTileThreads) {
      int4 rr = P.in[j];
      if (rr.w == dst && rr.x == slot && ((P.accept_bits[j >> 5] >> (j & 31)) & 1u)) atomicMax(&s_win, j);
    }
    __syncthreads();
    if (tid == 0 && s_win >= 0) {
      int4 rr = P.in[s_win];
      int l = local_slot(g, slot);
      int v = voter_index(g, dst >> 16, dst & 0xffff, slot);
      P.votes[(size_t)l * g.voters + v] = ((unsigned long long)(uint32_t)(rr.y + 1) << 32) | (uint32_t)rr.z;
    }
    __syncthreads();
  }
  if (tid == 0) P.st->n_conflicts = 0;
}

// ===========================================================================
// K3/K4  ProxyLeader.handlePhase2b   S/multipaxos/ProxyLeader.scala:217-258
//   The reference applies votes one at a time: phase2bs((g,a)) = msg (:237,
//   idempotent per acceptor), then the quorum test (:238-243); the FIRST vote
//   that makes the test pass sends Chosen (:246-253) and flips the key to Done
//   (:256); later votes see Done (:227-232).  Which vote completes depends on
//   delivery order, and the order of Chosen records in the output is the order
//   of their completing votes.  Two passes over the batch, no sort:
//     K3 stamp:    stamp[key][voter] = min(stamp, seq_i)      one RED per record
//     K4 complete: record i is the completing vote of its key iff it is the
//                  first delivery of its voter (stamp == seq_i), the voters with
//                  stamp < seq_i are NOT a quorum, and with voter i they ARE.
//                  Completing records are compacted in index order with a
//                  decoupled look-back, which yields the Chosen stream exactly.
//   seq_i = seq_base + i is a per-engine running sequence number, so first
//   deliveries of earlier batches order before this batch.
// ===========================================================================
struct TallyParams {
  Geometry g;
  PLState pl;
  const int4* in;
  int32_t n;
  uint32_t seq_base;
  int2* out_chosen;
  unsigned long long* desc_cnt;
  uint32_t epoch;
  DevStatus* st;
};

__global__ void __launch_bounds__(256) tally_stamp_kernel(TallyParams P) {
  const Geometry& g = P.g;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  int4 rec = ld_stream(P.in + i);  // {group, acceptor, slot, round}
  int slot = rec.z, round = rec.w;
  int local = local_slot(g, slot);
  if (local < 0) { report_error(P.st, FPX_ERR_SLOT_RANGE, i); return; }
  RowRef r = find_row(g, P.pl, local, slot, round);
  if (r.p == nullptr) { report_error(P.st, FPX_ERR_UNKNOWN_SLOT_ROUND, i); return; }  // :220-225
  int v = voter_index(g, rec.x, rec.y, slot);
  if (v < 0) return;                              // judged in K4 (needs Done-ness at i)
  if (r.round_word() & kDoneBit) return;          // Done before this batch (:227-232)
  atomicMin(&r.stamps()[v], P.seq_base + (uint32_t)i);
}

template <int ROWW>
__global__ void __launch_bounds__(kTileThreads) tally_complete_kernel(TallyParams P) {
  const Geometry& g = P.g;
  __shared__ uint32_t s_cnt[kChunks];
  __shared__ uint32_t s_tile_excl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const long long tile_base = (long long)tile * kTile;

  bool complete[kTileItems];
  int2 out[kTileItems];
  uint32_t rank[kTileItems];
  uint32_t* done_word[kTileItems];

#pragma unroll
  for (int it = 0; it < kTileItems; ++it) {
    long long i = tile_base + it * kTileThreads + tid;
    complete[it] = false;
    done_word[it] = nullptr;
    out[it] = make_int2(0, 0);
    if (i < P.n) {
      int4 rec = ld_stream(P.in + i);
      int slot = rec.z, round = rec.w;
      int local = local_slot(g, slot);
      RowRef r = local >= 0 ? find_row(g, P.pl, local, slot, round) : RowRef{nullptr};
      if (r.p != nullptr) {
        uint32_t w[ROWW];
        const int4* rp = (const int4*)r.p;
#pragma unroll
        for (int q = 0; q < ROWW / 4; ++q) {
          int4 t = __ldcg(rp + q);
          w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
        const uint32_t seq = P.seq_base + (uint32_t)i;
        uint32_t before = 0;
#pragma unroll
        for (int v = 0; v < ROWW - 2; ++v)
          if (v < g.voters && w[2 + v] < seq) before |= 1u << v;
        const bool done_before = write_quorum(g, before);
        const int v = voter_index(g, rec.x, rec.y, slot);
        if (!done_before) {
          if (v < 0) {
            // Grid.isWriteQuorum `require(xs subsetOf nodes)` (Grid.scala:44-47)
            report_error(P.st, FPX_ERR_BAD_ACCEPTOR, i);
          } else {
            uint32_t mine = 0;
#pragma unroll
            for (int q = 0; q < ROWW - 2; ++q)
              if (q == v) mine = w[2 + q];
            if (mine == seq && write_quorum(g, before | (1u << v))) {
              complete[it] = true;
              out[it] = make_int2(slot, (int)w[1]);  // Chosen(slot, pending.phase2a.value) (:249-251)
              done_word[it] = r.p;
            }
          }
        }
      }
    }
    unsigned b = __ballot_sync(0xffffffffu, complete[it]);
    rank[it] = __popc(b & lanemask_lt());
    if (lane == 0) s_cnt[it * (kTileThreads / 32) + warp] = __popc(b);
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t cnt = s_cnt[lane];
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    s_cnt[lane] = incl - cnt;
    uint32_t tile_total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) {
      uint32_t excl = lookback_sum(P.desc_cnt, tile, P.epoch, tile_total);
      s_tile_excl = excl;
      if (tile == (int)gridDim.x - 1) P.st->n_chosen = (int)(excl + tile_total);
    }
  }
  __syncthreads();
  const uint32_t tile_excl = s_tile_excl;
#pragma unroll
  for (int it = 0; it < kTileItems; ++it) {
    if (complete[it]) {
      uint32_t pos = tile_excl + s_cnt[it * (kTileThreads / 32) + warp] + rank[it];
      st_stream2(P.out_chosen + pos, out[it]);
      atomicOr(done_word[it], kDoneBit);  // states(slotround) = Done (:256)
    }
  }
}

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

__global__ void __launch_bounds__(256) replica_chosen_kernel(ReplicaParams P) {
  const Geometry& g = P.g;
  int n = P.n >= 0 ? P.n : P.st->n_chosen;
  int mx = INT_MIN;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int2 rec = P.in[i];
    int local = local_slot(g, rec.x);
    if (local < 0) { report_error(P.st, FPX_ERR_SLOT_RANGE, i); continue; }
    unsigned long long w = ((unsigned long long)(P.seq_base + (uint32_t)i) << 32) | (uint32_t)rec.y;
    atomicMin(&P.rlog[local], w);
    mx = max(mx, local);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx != INT_MIN) atomicMax(&P.st->max_chosen_local, mx);
}

__global__ void __launch_bounds__(256) watermark_scan_kernel(Geometry g, const unsigned long long* rlog,
                                                            DevStatus* st) {
  int lo = st->wm_local;
  int hi = min(st->max_chosen_local + 2, g.local_slots);  // one past the last candidate hole
  for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
    if (i >= *(volatile int*)&st->wm_found) break;
    if (rlog[i] == kU64Empty) { atomicMin(&st->wm_found, i); break; }
  }
}
__global__ void watermark_finish_kernel(Geometry g, DevStatus* st, int32_t* d_out) {
  int hi = min(st->max_chosen_local + 2, g.local_slots);
  int found = min(st->wm_found, hi);
  found = max(found, st->wm_local);
  if (found > g.local_slots) found = g.local_slots;
  st->wm_local = found;
  st->wm_found = INT_MAX;
  int global = found * g.shard_count + g.shard_index;
  st->watermark = global;
  if (d_out) *d_out = global;
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
  int vr = -1, vv = -1;
  if (v >= 0) {
    unsigned long long c = votes[(size_t)l * g.voters + v];
    if ((c >> 32) != 0) { vr = (int)(c >> 32) - 1; vv = (int)(uint32_t)c; }
  }
  vote_round[i] = vr;
  vote_value[i] = vv;
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
