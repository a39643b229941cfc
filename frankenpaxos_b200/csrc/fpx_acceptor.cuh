// fpx_acceptor.cuh -- K2: Acceptor.handlePhase2a for the interleaved delivery
// stream of all acceptors of the config.   S/multipaxos/Acceptor.scala:184-220
//
//   `round` is ONE scalar per acceptor (:95), so for record i addressed to
//   acceptor k the handler's test (:192) is
//        msg.round < max(round_k at batch start, max_{j<i, dst_j==k} msg_j.round)
//   (rejected messages are below the running max, so including them is
//   harmless): an exclusive keyed prefix-max in delivery order.
//
//   Persistent cooperative kernel, every warp owns a contiguous range:
//     pass 1  stream the range once, lane k of the warp accumulates the max round
//             addressed to acceptor k (warp redux; one instruction when the whole
//             chunk carries one round -- the steady state);
//     barrier CTA aggregates -> global; every CTA takes the max over the CTAs
//             before it (+ the acceptors' rounds at batch start) = its carry-in;
//     pass 2  stream the range again (now an L2 hit): accept test, vote cell,
//             Phase2b reply, maxVotedSlot.  Replies are written at index i
//             (dense = correct whenever the batch produces no Nack);
//     barrier if any Nack was produced (leader change): pass 3 rewrites both
//             reply streams compacted in delivery order from exact prefix counts.
//   Vote cell: states(slot) = State(round, value) (:205-208) is one 64-bit
//   atomicMax of (round+1 : value_id): accepted rounds never decrease in
//   delivery order, so max == last writer, except "same round, different value"
//   which is flagged and resolved to last-in-order by the last block.
#pragma once
#include "fpx_common.cuh"

namespace fpx {

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
  int32_t* g_agg;                  // [grid][kMaxKeys] per-CTA max round per acceptor
  uint32_t* g_wacc;                // [grid*kAW] accepted records per warp range
  uint32_t parity;                 // which nack counter this launch uses
  int32_t append;                  // 1: continue the reply streams of the previous launch (chunked host call)
  DevStatus* st;
  VoteConflict* conflicts;
};

constexpr int kAccUnroll = 4;
// one CTA of 32 warps per SM (see fpx_tally.cuh: several CTAs per SM spread up to 2x, and both passes end at a grid barrier)
#ifndef FPX_AT
#define FPX_AT 1024
#endif
constexpr int kAT = FPX_AT;
constexpr int kAW = kAT / 32;
__device__ __forceinline__ int grp_of(int dst) { return dst >> 16; }

// decode + validate one Phase2a record; key = global acceptor id or -1
__device__ __forceinline__ int decode_p2a(const Geometry& g, const int4& rec, int& key, int& loc, int& vix) {
  key = -1; loc = -1; vix = -1;
  int grp = rec.w >> 16, acc = rec.w & 0xffff;
  if ((uint32_t)grp >= (uint32_t)g.groups || acc >= g.per_group) return FPX_ERR_BAD_ACCEPTOR;
  if ((uint32_t)rec.y > (uint32_t)FPX_MAX_ROUND) return FPX_ERR_ROUND_RANGE;
  int l = local_slot(g, rec.x);
  if (l < 0 && l != kLocalRetired) return FPX_ERR_SLOT_RANGE;   // retired: round compare and reply as usual, the vote cell is gone
  int v = voter_index(g, grp, acc, rec.x);
  if (v < 0 || (g.protocol == FPX_MENCIUS && grp != expected_group(g, rec.x))) return FPX_ERR_BAD_ACCEPTOR;
  key = grp * g.per_group + acc; loc = l; vix = v;
  return 0;
}
// pass 1 only needs the acceptor id and the round; slot-level validity is judged
// (and reported) in pass 2
__device__ __forceinline__ int decode_key(const Geometry& g, const int4& rec) {
  int grp = rec.w >> 16, acc = rec.w & 0xffff;
  if ((uint32_t)grp >= (uint32_t)g.groups || acc >= g.per_group) return -1;
  if ((uint32_t)rec.y > (uint32_t)FPX_MAX_ROUND) return -1;
  return grp * g.per_group + acc;
}
__device__ __noinline__ void acceptor_error(DevStatus* st, int code, int index) { report_error(st, code, index); }

// Pass 2 (kExact = false: effects + dense replies) and pass 3 (kExact = true:
// compacted replies only).  `run` is lane-indexed: lane k holds acceptor k's
// round as of the start of the warp's range.  s_mv is a [num_keys][kAT]
// table of per-thread private maxima of accepted slots (maxVotedSlot, :209).
template <bool kExact>
__device__ __forceinline__ void acceptor_apply(const AcceptorParams& P, int4* out_p2b, int2* out_nack, int wlo, int whi,
                                               int lane, int run, uint32_t pos_base, int* s_mv, uint32_t& wacc,
                                               uint32_t& wnack) {
  const Geometry& g = P.g;
  const unsigned full = 0xffffffffu;
  // the reply stream is write-once: evict_first gets it written back while this kernel
  // still runs instead of during the next kernel's reads
  const unsigned long long pol_out = l2_policy_evict_first();
  for (int base = wlo; base < whi; base += 32 * kAccUnroll) {
    int4 rec[kAccUnroll];
    unsigned long long cell[kAccUnroll], old[kAccUnroll];
#pragma unroll
    for (int u = 0; u < kAccUnroll; ++u) {
      int i = base + u * 32 + lane;
      rec[u] = (i < whi) ? ld_cg(P.in + i) : make_int4(0, -1, 0, -1);
      cell[u] = 0; old[u] = 0;
    }
#pragma unroll
    for (int u = 0; u < kAccUnroll; ++u) {
      const int i0 = base + u * 32;
      if (i0 >= whi) break;
      const int i = i0 + lane;
      int key, loc, vix;
      bool valid = false;
      if (i < whi) {
        int err = decode_p2a(g, rec[u], key, loc, vix);
        if (err && !kExact) acceptor_error(P.st, err, i);
        valid = err == 0;
      }
      const int r = rec[u].y;
      int mn = __reduce_min_sync(full, valid ? r : INT_MAX);
      int mx = __reduce_max_sync(full, valid ? r : INT_MIN);
      if (mx == INT_MIN) {
        if (!kExact && lane == 0) P.accept_bits[i0 >> 5] = 0;
        continue;
      }
      int my_run = __shfl_sync(full, run, key & 31);
      int cur;
      if (mn == mx) {  // one round in the whole chunk: no in-chunk dependency
        cur = my_run;
        unsigned present = __reduce_or_sync(full, valid ? (1u << key) : 0u);
        if ((present >> lane) & 1u) run = max(run, mx);
      } else {
        int pin = INT_MIN, chunk_agg = INT_MIN;
        unsigned remaining = __ballot_sync(full, valid);
        while (remaining) {
          int leader = __ffs(remaining) - 1;
          int kk = __shfl_sync(full, key, leader);
          bool mine = valid && key == kk;
          int incl = warp_incl_scan_max(mine ? r : INT_MIN, lane);
          int ex = __shfl_up_sync(full, incl, 1);
          if (lane == 0) ex = INT_MIN;
          if (mine) pin = ex;
          int tot = __shfl_sync(full, incl, 31);
          if (lane == kk) chunk_agg = tot;
          remaining &= ~__ballot_sync(full, mine);
        }
        cur = max(my_run, pin);
        run = max(run, chunk_agg);
      }
      const bool accept = valid && r >= cur;   // Acceptor.scala:192
      const unsigned b = __ballot_sync(full, accept);
      if (!kExact) {
        if (lane == 0) P.accept_bits[i0 >> 5] = b;
        if (accept) {
          // Phase2b(groupIndex, acceptorIndex, slot, round) (:211-219)
          st_evict_first(out_p2b + i, make_int4(rec[u].w >> 16, rec[u].w & 0xffff, rec[u].x, r), pol_out);
          // states(slot) = State(voteRound = round, voteValue) (:205-208)
          if (loc >= 0) {   // (a retired slot's vote is never read again: Phase1b starts at the chosen watermark, :171-179)
            cell[u] = ((unsigned long long)(uint32_t)(r + 1) << 32) | (uint32_t)rec[u].z;
            old[u] = atomicMax(&P.votes[cell_index(g, loc, vix)], cell[u]);
          }
          // maxVotedSlot = max(maxVotedSlot, slot) (:209): thread-private column
          atomicMax(&s_mv[key * kAT + threadIdx.x], rec[u].x);
        }
        wnack += __popc(__ballot_sync(full, valid && !accept));
      } else {
        uint32_t before = pos_base + wacc + __popc(b & lanemask_lt());
        if (accept) {
          st_stream(out_p2b + before, make_int4(rec[u].w >> 16, rec[u].w & 0xffff, rec[u].x, r));
        } else if (valid) {
          // Nack(round) to leaders(roundSystem.leader(phase2a.round)) (:197-198)
          int ldr = r % g.num_leaders;
          if (g.protocol == FPX_MENCIUS) ldr += (grp_of(rec[u].w) / g.agroups) * g.num_leaders;
          st_stream2(out_nack + ((uint32_t)i - before), make_int2(ldr, cur));
        }
      }
      wacc += __popc(b);
    }
    if (!kExact) {
      // the atomics' return values are only consumed here, after all of the
      // group's loads/atomics are in flight
#pragma unroll
      for (int u = 0; u < kAccUnroll; ++u) {
        if ((old[u] >> 32) == (cell[u] >> 32) && old[u] != cell[u]) {
          uint32_t cidx = atomicAdd(&P.st->n_conflicts, 1u);
          if (cidx < (uint32_t)kMaxConflicts) P.conflicts[cidx] = VoteConflict{rec[u].w, rec[u].x};
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kAT, 1024 / kAT) acceptor_phase2a_kernel(AcceptorParams P) {
  const Geometry& g = P.g;
  extern __shared__ int s_mv[];  // [num_keys][kAT] private maxVotedSlot columns
  __shared__ int s_wagg[kAW][kMaxKeys];
  __shared__ int s_tmp[kAW][kMaxKeys];
  __shared__ int s_win;

  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total_warps = gridDim.x * kAW;
  const int per = (((P.n + total_warps - 1) / total_warps) + 31) & ~31;   // contiguous range of every warp
  const int gw = blockIdx.x * kAW + warp;
  const int wlo = (int)min((long long)P.n, (long long)gw * per);
  const int whi = (int)min((long long)P.n, (long long)wlo + per);
  uint32_t* nack_ctr = P.parity ? &P.st->nack_total : &P.st->pad[0];
  uint32_t* nack_other = P.parity ? &P.st->pad[0] : &P.st->nack_total;
  if (blockIdx.x == 0 && tid == 0) *nack_other = 0;  // the counter the NEXT launch uses
  // reply-stream bases: 0, or where the previous launch of a chunked call stopped (read by
  // every CTA before the first grid barrier; rewritten only after the second)
  const uint32_t base_p2b = P.append ? (uint32_t)__ldcg(&P.st->n_p2b) : 0u;
  const uint32_t base_nack = P.append ? (uint32_t)__ldcg(&P.st->n_nack) : 0u;
  int4* const out_p2b = P.out_p2b + base_p2b;
  int2* const out_nack = P.out_nack + base_nack;
  for (int k = 0; k < g.num_keys; ++k) s_mv[k * kAT + tid] = INT_MIN;
  FPX_MARK(P.st->t_acceptor, 0);

  // ---- pass 1: per-acceptor max round of the warp's range (lane = acceptor).
  // The stream is tagged evict_last in L2 so that pass 2 (which also writes
  // 88 B/record of replies and vote cells through L2) still finds it there.
  const unsigned long long pol_keep = l2_policy_evict_last();
  int wagg = INT_MIN;
  for (int base = wlo; base < whi; base += 32 * kAccUnroll) {
    int4 rec[kAccUnroll];
#pragma unroll
    for (int u = 0; u < kAccUnroll; ++u) {
      int i = base + u * 32 + lane;
      rec[u] = (i < whi) ? ld_keep(P.in + i, pol_keep) : make_int4(0, -1, 0, -1);
    }
#pragma unroll
    for (int u = 0; u < kAccUnroll; ++u) {
      if (base + u * 32 >= whi) break;
      const int key = decode_key(g, rec[u]);  // -1 for padding lanes too (round -1)
      const bool valid = key >= 0;
      const int r = rec[u].y;
      int mn = __reduce_min_sync(full, valid ? r : INT_MAX);
      int mx = __reduce_max_sync(full, valid ? r : INT_MIN);
      if (mx == INT_MIN) continue;
      if (mn == mx) {
        unsigned present = __reduce_or_sync(full, valid ? (1u << key) : 0u);
        if ((present >> lane) & 1u) wagg = max(wagg, mx);
      } else {
        unsigned remaining = __ballot_sync(full, valid);
        while (remaining) {
          int leader = __ffs(remaining) - 1;
          int kk = __shfl_sync(full, key, leader);
          bool mine = valid && key == kk;
          int m = __reduce_max_sync(full, mine ? r : INT_MIN);
          if (lane == kk) wagg = max(wagg, m);
          remaining &= ~__ballot_sync(full, mine);
        }
      }
    }
  }
  s_wagg[warp][lane] = wagg;
  __syncthreads();
  int cta_agg = INT_MIN;
  if (warp == 0) {
#pragma unroll
    for (int w = 0; w < kAW; ++w) cta_agg = max(cta_agg, s_wagg[w][lane]);
    __stcg(&P.g_agg[blockIdx.x * kMaxKeys + lane], cta_agg);
  }
  FPX_MARK(P.st->t_acceptor, 1);
  grid_sync(P.st);
  FPX_MARK(P.st->t_acceptor, 2);

  // ---- carry-in: acceptor rounds at batch start + every CTA before this one.
  // thread t covers CTAs t, t+256, ... for every acceptor: all loads independent.
  for (int k = 0; k < g.num_keys; ++k) {
    int v = INT_MIN;
    for (int c = tid; c < (int)blockIdx.x; c += kAT) v = max(v, __ldcg(&P.g_agg[c * kMaxKeys + k]));
    v = __reduce_max_sync(full, v);
    if (lane == 0) s_tmp[warp][k] = v;
  }
  __syncthreads();
  int cta_carry = INT_MIN;
  if (lane < g.num_keys) {
    cta_carry = __ldcg(&P.acc_round[lane]);
#pragma unroll
    for (int w = 0; w < kAW; ++w) cta_carry = max(cta_carry, s_tmp[w][lane]);
  }
  int run = cta_carry;
  for (int w = 0; w < warp; ++w) run = max(run, s_wagg[w][lane]);

  FPX_MARK(P.st->t_acceptor, 3);
  // ---- pass 2: decisions + effects, replies at dense positions
  uint32_t wacc = 0, wnack = 0;
  acceptor_apply<false>(P, out_p2b, out_nack, wlo, whi, lane, run, 0u, s_mv, wacc, wnack);
  if (lane == 0) {
    __stcg(&P.g_wacc[gw], wacc);
    if (wnack) atomicAdd(nack_ctr, wnack);
  }
  __syncthreads();
  for (int k = warp; k < g.num_keys; k += kAW) {
    int m = INT_MIN;
#pragma unroll
    for (int t = lane; t < kAT; t += 32) m = max(m, s_mv[k * kAT + t]);
    m = __reduce_max_sync(full, m);
    if (lane == 0 && m != INT_MIN) atomicMax(&P.acc_max_voted[k], m);
  }
  FPX_MARK(P.st->t_acceptor, 4);
  grid_sync(P.st);
  FPX_MARK(P.st->t_acceptor, 5);

  // round after the batch = max over everything (:204); only now is it safe to
  // overwrite the batch-start value every CTA read above
  if (blockIdx.x == gridDim.x - 1 && warp == 0 && lane < g.num_keys) {
    int tot = cta_carry;
#pragma unroll
    for (int w = 0; w < kAW; ++w) tot = max(tot, s_wagg[w][lane]);
    P.acc_round[lane] = tot;
  }
  const uint32_t total_nacks = __ldcg(nack_ctr);
  if (total_nacks == 0) {
    if (blockIdx.x == 0 && tid == 0) { P.st->n_p2b = (int)base_p2b + P.n; P.st->n_nack = (int)base_nack; }
  } else {
    // ---- pass 3 (leader change only): exact, compacted reply streams
    uint32_t before = 0;
    for (int j = lane; j < gw; j += 32) before += __ldcg(&P.g_wacc[j]);
    before = __reduce_add_sync(full, before);
    uint32_t wacc2 = 0, wnack2 = 0;
    acceptor_apply<true>(P, out_p2b, out_nack, wlo, whi, lane, run, before, s_mv, wacc2, wnack2);
    if (gw == (int)gridDim.x * kAW - 1 && lane == 0) {
      P.st->n_p2b = (int)(base_p2b + before + wacc2);
      P.st->n_nack = (int)base_nack + P.n - (int)(before + wacc2);
    }
  }

  // ---- CTA 0: same (acceptor, slot, round) voted twice with different values in one batch -> the
  // later delivery must win (map overwrite, :205).  Every conflict was recorded in pass 2, i.e. before
  // the second grid barrier.
  if (blockIdx.x != 0) return;
  uint32_t nc = __ldcg(&P.st->n_conflicts);
  if (nc == 0) return;
  if (nc > (uint32_t)kMaxConflicts) {
    if (tid == 0) { report_error(P.st, FPX_ERR_CONFLICT, 0); P.st->n_conflicts = 0; }
    return;
  }
  for (uint32_t cix = 0; cix < nc; ++cix) {
    int dst = P.conflicts[cix].dst, slot = P.conflicts[cix].slot;
    if (tid == 0) s_win = -1;
    __syncthreads();
    for (int j = tid; j < P.n; j += kAT) {
      int4 rr = P.in[j];
      if (rr.w == dst && rr.x == slot && ((__ldcg(&P.accept_bits[j >> 5]) >> (j & 31)) & 1u)) atomicMax(&s_win, j);
    }
    __syncthreads();
    if (tid == 0 && s_win >= 0) {
      int4 rr = P.in[s_win];
      int l = local_slot(g, slot);
      int v = voter_index(g, dst >> 16, dst & 0xffff, slot);
      P.votes[cell_index(g, l, v)] = ((unsigned long long)(uint32_t)(rr.y + 1) << 32) | (uint32_t)rr.z;
    }
    __syncthreads();
  }
  if (tid == 0) P.st->n_conflicts = 0;
}

}  // namespace fpx
