// fpx_tally.cuh -- K3: ProxyLeader.handlePhase2b   S/multipaxos/ProxyLeader.scala:217-258
//
//   The reference applies votes one at a time: phase2bs((g,a)) = msg (:237,
//   idempotent per acceptor), then the quorum test (:238-243); the FIRST vote
//   that makes the test pass sends Chosen (:246-253) and flips the key to Done
//   (:256); later votes see Done (:227-232).  Which vote completes depends on
//   delivery order, and the order of Chosen records in the output is the order
//   of their completing votes.  One persistent cooperative kernel, every warp
//   owns a contiguous range of the delivery stream, no sort:
//     phase A  stamp[slot][voter] = min(stamp, seq_i)         one RED per record
//     barrier
//     phase B  record i is the completing vote of its key iff it is the first
//              delivery of its voter (stamp == seq_i), the voters with
//              stamp < seq_i are NOT a quorum, and with voter i they ARE.  The
//              warp appends its completing records, in order, to its slice of a
//              shared-memory buffer;
//     barrier  CTA counts -> global, each CTA sums the CTAs before it
//     phase C  coalesced copy of the buffered Chosen records to their exact
//              positions: the output is the Chosen stream in delivery order.
//   seq_i = seq_base + i is a per-engine running sequence number, so first
//   deliveries of earlier batches order before this batch.  `Done` (:256) is not
//   stored: a key is Done at delivery i iff the stamps below seq_i already form
//   a quorum, which is exactly what phase B evaluates (later votes of a Done key
//   only add larger stamps, which never change that).
//
//   Cost model (measured, profiles/): the Phase2b stream is shuffled, so every
//   row access is a fully divergent warp access = one L1TEX wavefront per lane
//   PER INSTRUCTION (~23 us per divergent instruction per 3*2^20 votes).  A row
//   {round, value, stamp[voters]} is ONE 32-byte sector for <= 6 voters, read
//   with one 256-bit load; and phase A stamps BLINDLY -- without loading the
//   row's header to check the round -- which is sound because a slot whose
//   primary row is in normal state has exactly ONE armed round: arming a second
//   round of a slot (leader change) moves the first one to the (slot, round) table
//   and poisons the primary row (fpx_arm.cuh), after which the row's stamps are
//   never read again.  Phase B reads the header anyway; a vote that finds a poisoned
//   header is stamped into its table entry there, and if any vote did, all warps
//   redo phase B after the barrier (leader-change batches only).  Vanilla Mencius
//   ignores stale-round votes instead of failing on them, so it keeps the checked
//   (header-loading) phase A.
#pragma once
#include "fpx_common.cuh"

namespace fpx {

struct TallyParams {
  Geometry g;
  PLState pl;
  const int4* in;
  int32_t n;
  uint32_t seq_base;
  int2* out_chosen;
  uint32_t* g_ccnt;        // [grid] Chosen records produced per CTA
  uint32_t bar_base;
  int32_t first;           // 1: first launch of a call (output offset 0), else append at st->n_chosen
  int32_t per;             // records per warp range == shared buffer entries per warp
  unsigned long long* votes;  // vanilla Mencius: the coordinator's log entry turns ChosenEntry on completion
  DevStatus* st;
};

constexpr int kTallyUnroll = 4;  // chunks in flight per warp (phase B: 32 / ROWW)

// One proxy-leader row from L2 with a single 256-bit load per 8 words (LDG.E.256, sm_100+).
template <int ROWW>
__device__ __forceinline__ void load_row(const uint32_t* p, uint32_t (&w)[ROWW]) {
#pragma unroll
  for (int q = 0; q < ROWW / 8; ++q) {
    asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[8 * q]), "=r"(w[8 * q + 1]), "=r"(w[8 * q + 2]), "=r"(w[8 * q + 3]),
                   "=r"(w[8 * q + 4]), "=r"(w[8 * q + 5]), "=r"(w[8 * q + 6]), "=r"(w[8 * q + 7])
                 : "l"(p + 8 * q));
  }
}

// Phase B over the warp's range.  kRedo = false: first evaluation; a vote whose slot
// is poisoned is stamped into its table entry and the any_slow flag is raised.
// kRedo = true: everything is evaluated, poisoned slots against their (now fully
// stamped) table entry.
template <int ROWW, bool kRedo>
__device__ __forceinline__ uint32_t tally_phase_b(const TallyParams& P, int wlo, int whi, int lane, int2* my_buf) {
  const Geometry& g = P.g;
  const unsigned full = 0xffffffffu;
  const bool vanilla = g.protocol == FPX_VANILLA_MENCIUS;
  constexpr int kUB = ROWW <= 16 ? 2 : 1;
  uint32_t wcnt = 0;
  for (int base = wlo; base < whi; base += 32 * kUB) {
    int4 rec[kUB];
#pragma unroll
    for (int u = 0; u < kUB; ++u) {
      int i = base + u * 32 + lane;
      rec[u] = (i < whi) ? ld_cg(P.in + i) : make_int4(-1, -1, -1, -1);  // {group, acceptor, slot, round}
    }
    uint32_t w[kUB][ROWW];
    uint32_t* row[kUB];
#pragma unroll
    for (int u = 0; u < kUB; ++u) {
      int i = base + u * 32 + lane;
      row[u] = nullptr;
      if (i < whi) {
        int local = local_slot(g, rec[u].z);
        if (local >= 0) {
          row[u] = P.pl.rows + (size_t)local * g.row_words;
          load_row<ROWW>(row[u], w[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUB; ++u) {
      if (base + u * 32 >= whi) break;
      const int i = base + u * 32 + lane;
      bool complete = false;
      int2 out = make_int2(0, 0);
      if (row[u] != nullptr) {
        uint32_t* r = row[u];
        const uint32_t hw = w[u][0];
        bool ok = true;
        if (vanilla) {
          // Server.handlePhase2b: no Phase 2 for the slot / already chosen -> ignore
          // (:1088-1106); stale round -> ignore (:1109-1112); larger: checkEq (:1116), phase A
          ok = hw != kUnarmed && !(hw & kDoneBit) && rec[u].w == (int)hw;
        } else if (hw == kPoison) {
          // several rounds of this slot were armed: the key lives in the table
          RowRef rr = find_row(g, P.pl, local_slot(g, rec[u].z), rec[u].z, rec[u].w);
          if (rr.p == nullptr) {
            if (!kRedo) report_error(P.st, FPX_ERR_UNKNOWN_SLOT_ROUND, i);            // :220-225
            ok = false;
          } else if (!kRedo) {
            const int v = voter_index(g, rec[u].x, rec[u].y, rec[u].z);
            if (v >= 0) atomicMin(&rr.p[2 + v], P.seq_base + (uint32_t)i);            // late stamp (:237)
            P.st->pad[1] = 1;                                                          // any_slow
            ok = false;
          } else {
            r = rr.p;
            load_row<ROWW>(r, w[u]);
          }
        } else if (hw == kUnarmed || (int)hw != rec[u].w) {
          // the slot has one armed round and it is not this vote's: never armed (:220-225)
          if (!kRedo) report_error(P.st, FPX_ERR_UNKNOWN_SLOT_ROUND, i);
          ok = false;
        }
        if (ok) {
          const uint32_t seq = P.seq_base + (uint32_t)i;
          uint32_t before = 0;
#pragma unroll
          for (int v = 0; v < ROWW - 2; ++v)
            if (v < g.voters && w[u][2 + v] < seq) before |= 1u << v;
          if (!write_quorum(g, before)) {  // key still Pending when vote i is delivered
            const int v = voter_index(g, rec[u].x, rec[u].y, rec[u].z);
            if (v < 0) {
              // Grid.isWriteQuorum `require(xs subsetOf nodes)` (Grid.scala:44-47)
              if (!kRedo || hw == kPoison) report_error(P.st, FPX_ERR_BAD_ACCEPTOR, i);
            } else {
              uint32_t mine = 0;
#pragma unroll
              for (int q = 0; q < ROWW - 2; ++q)
                if (q == v) mine = w[u][2 + q];
              if (mine == seq && write_quorum(g, before | (1u << v))) {
                complete = true;
                out = make_int2(rec[u].z, (int)w[u][1]);  // Chosen(slot, pending.phase2a.value) (:249-251)
                if (vanilla) {
                  // choose(): the coordinator's own entry becomes ChosenEntry, phase2s.remove (:622-625)
                  int owner = rec[u].z % g.per_group;
                  atomicMax(&P.votes[(size_t)local_slot(g, rec[u].z) * g.voters + owner], kCellChosen | w[u][1]);
                  atomicOr(r, kDoneBit);
                }
              }
            }
          }
        }
      }
      unsigned b = __ballot_sync(full, complete);
      if (complete) my_buf[wcnt + __popc(b & lanemask_lt())] = out;
      wcnt += __popc(b);
    }
  }
  return wcnt;
}

template <int ROWW>
__global__ void __launch_bounds__(kThreads, ROWW == 8 ? 4 : 1) tally_kernel(TallyParams P) {
  const Geometry& g = P.g;
  extern __shared__ int2 s_buf[];  // kWarps * P.per Chosen records
  __shared__ uint32_t s_wcnt[kWarps];
  __shared__ uint32_t s_woff[kWarps];
  __shared__ uint32_t s_cta_off;

  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = P.per;
  const int gw = blockIdx.x * kWarps + warp;
  const int wlo = (int)min((long long)P.n, (long long)gw * per);
  const int whi = (int)min((long long)P.n, (long long)wlo + per);
  const bool vanilla = g.protocol == FPX_VANILLA_MENCIUS;

  // ---- phase A: first-delivery stamps (stream tagged evict_last: phase B re-reads it from L2)
  const unsigned long long pol_keep = l2_policy_evict_last();
  FPX_MARK(P.st->t_tally, 0);
  for (int base = wlo; base < whi; base += 32 * kTallyUnroll) {
    int4 rec[kTallyUnroll];
#pragma unroll
    for (int u = 0; u < kTallyUnroll; ++u) {
      int i = base + u * 32 + lane;
      rec[u] = (i < whi) ? ld_keep(P.in + i, pol_keep) : make_int4(-1, -1, -1, -1);  // {group, acceptor, slot, round}
    }
    if (!vanilla) {
      // blind: no row header load (see the file comment); phase2bs((g,a)) = msg (:237)
#pragma unroll
      for (int u = 0; u < kTallyUnroll; ++u) {
        int i = base + u * 32 + lane;
        if (i >= whi) continue;
        int local = local_slot(g, rec[u].z);
        if (local < 0) { report_error(P.st, FPX_ERR_SLOT_RANGE, i); continue; }
        int v = voter_index(g, rec[u].x, rec[u].y, rec[u].z);
        if (v < 0) continue;                                         // judged in phase B (needs Done-ness at i)
        atomicMin(P.pl.rows + (size_t)local * g.row_words + 2 + v, P.seq_base + (uint32_t)i);
      }
    } else {
      uint32_t* row[kTallyUnroll];
      uint32_t rw[kTallyUnroll];
#pragma unroll
      for (int u = 0; u < kTallyUnroll; ++u) {
        int i = base + u * 32 + lane;
        row[u] = nullptr;
        rw[u] = kUnarmed;
        if (i < whi) {
          int local = local_slot(g, rec[u].z);
          if (local < 0) {
            report_error(P.st, FPX_ERR_SLOT_RANGE, i);
          } else {
            row[u] = P.pl.rows + (size_t)local * g.row_words;
            rw[u] = __ldcg(row[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kTallyUnroll; ++u) {
        int i = base + u * 32 + lane;
        if (row[u] == nullptr) continue;
        const uint32_t w = rw[u];
        // Server.handlePhase2b: no Phase 2 for the slot / already chosen -> ignore
        // (:1088-1106); stale round -> ignore (:1109-1112); a larger round fails checkEq (:1116)
        if (w == kUnarmed || (w & kDoneBit) || rec[u].w < (int)w) continue;
        if (rec[u].w > (int)w) { report_error(P.st, FPX_ERR_UNKNOWN_SLOT_ROUND, i); continue; }
        int v = voter_index(g, rec[u].x, rec[u].y, rec[u].z);
        if (v < 0) continue;
        atomicMin(&row[u][2 + v], P.seq_base + (uint32_t)i);
      }
    }
  }
  FPX_MARK(P.st->t_tally, 1);
  grid_barrier(&P.st->barrier, P.bar_base + gridDim.x);
  FPX_MARK(P.st->t_tally, 2);

  // ---- phase B: completing votes, buffered in delivery order per warp
  int2* my_buf = s_buf + (size_t)warp * per;
  uint32_t wcnt = tally_phase_b<ROWW, false>(P, wlo, whi, lane, my_buf);
  auto publish = [&]() {
    if (lane == 0) s_wcnt[warp] = wcnt;
    __syncthreads();
    if (tid == 0) {
      uint32_t run = 0;
#pragma unroll
      for (int wv = 0; wv < kWarps; ++wv) { s_woff[wv] = run; run += s_wcnt[wv]; }
      __stcg(&P.g_ccnt[blockIdx.x], run);
    }
  };
  publish();
  FPX_MARK(P.st->t_tally, 3);
  grid_barrier(&P.st->barrier, P.bar_base + 2 * gridDim.x);
  FPX_MARK(P.st->t_tally, 4);
  if (__ldcg(&P.st->pad[1]) != 0) {
    // some votes belong to slots with several armed rounds (leader change): their table
    // entries are fully stamped now; evaluate everything again, in order
    wcnt = tally_phase_b<ROWW, true>(P, wlo, whi, lane, my_buf);
    __syncthreads();
    publish();
    grid_barrier(&P.st->barrier, P.bar_base + 3 * gridDim.x);
  } else if (tid == 0) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&P.st->barrier) : "memory");  // arrivals stay 4 per CTA
  }

  // ---- phase C: exact output positions, coalesced copy-out
  if (warp == 0) {
    uint32_t off = 0;
    for (int c = lane; c < (int)blockIdx.x; c += 32) off += __ldcg(&P.g_ccnt[c]);
    off = __reduce_add_sync(full, off);
    uint32_t out_base = P.first ? 0u : (uint32_t)__ldcg(&P.st->n_chosen);
    if (lane == 0) s_cta_off = out_base + off;
  }
  __syncthreads();
  {
    const uint32_t dst0 = s_cta_off + s_woff[warp];
    for (uint32_t j = lane; j < wcnt; j += 32) st_stream2(P.out_chosen + dst0 + j, my_buf[j]);
  }
  FPX_MARK(P.st->t_tally, 5);
  if (blockIdx.x == gridDim.x - 1) {
    // every CTA has read the old n_chosen / any_slow before this one may overwrite them
    grid_barrier(&P.st->barrier, P.bar_base + 4 * gridDim.x);
    if (tid == 0) {
      P.st->n_chosen = (int)(s_cta_off + s_woff[kWarps - 1] + s_wcnt[kWarps - 1]);
      P.st->pad[1] = 0;
    }
  } else {
    if (tid == 0) {
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&P.st->barrier) : "memory");
    }
  }
}

}  // namespace fpx
