// fpx_vanilla.cuh -- vanilla Mencius (S/vanillamencius/Server.scala), SURVEY 8(a) row a9.
// The per-slot round compare of Server.handlePhase2a (:1001-1082) is per log
// entry, so records of different (slot, server) cells are independent: one
// thread per record, one 64-bit atomicMax per vote cell.  Cell = {hi: round+1 |
// chosen bit 31, lo: value}; a ChosenEntry compares above every vote, so it is
// absorbing under atomicMax exactly like `case Some(chosen: ChosenEntry)` (:1017).
#pragma once
#include "fpx_common.cuh"

namespace fpx {

struct VmParams {
  Geometry g;
  const int4* in;
  int4* out;
  int32_t n;
  unsigned long long* votes;   // [local_slots][n]
  unsigned long long* claim;   // [local_slots][n]  (~batch tag : min index) -- one Phase2a per cell per batch
  uint32_t* rows;              // proxy-leader rows (phase2s)
  uint32_t tag;
  DevStatus* st;
};

// handlePhase2a.  in {slot, round, value, dst = server}; out {kind, server, slot, round|value}
__global__ void __launch_bounds__(256) vm_phase2a_kernel(VmParams P) {
  const Geometry& g = P.g;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  int4 rec = ld_stream(P.in + i);
  int server = rec.w & 0xffff;
  int4 rep = make_int4(-1, server, rec.x, 0);
  int local = local_slot(g, rec.x);
  if (local < 0) { report_error(P.st, FPX_ERR_SLOT_RANGE, i); P.out[i] = rep; return; }
  if ((rec.w >> 16) != 0 || server >= g.per_group) { report_error(P.st, FPX_ERR_BAD_ACCEPTOR, i); P.out[i] = rep; return; }
  if ((uint32_t)rec.y > (uint32_t)FPX_MAX_ROUND) { report_error(P.st, FPX_ERR_ROUND_RANGE, i); P.out[i] = rep; return; }
  size_t cellix = (size_t)local * g.voters + server;
  unsigned long long mine = ((unsigned long long)(~P.tag) << 32) | (uint32_t)i;
  unsigned long long oldc = atomicMin(&P.claim[cellix], mine);
  if ((uint32_t)(oldc >> 32) == ~P.tag) {
    report_error(P.st, FPX_ERR_BATCH_ORDER, max((long long)i, (long long)(uint32_t)oldc));
    P.out[i] = rep;
    return;
  }
  unsigned long long cell = ((unsigned long long)(uint32_t)(rec.y + 1) << 32) | (uint32_t)rec.z;
  unsigned long long old = atomicMax(&P.votes[cellix], cell);
  if (old & kCellChosen) {
    rep.x = 2; rep.w = (int)(uint32_t)old;                       // Chosen(slot, chosen.value) (:1018-1027)
  } else if ((uint32_t)(old >> 32) > (uint32_t)(cell >> 32)) {
    rep.x = 1; rep.w = (int)(uint32_t)(old >> 32) - 1;           // Phase2Nack(slot, round) (:1044-1051)
  } else {
    // log.put(slot, PendingEntry(round, round, value)) (:1054-1058); an equal round
    // re-vote overwrites: atomicMax kept the larger value, force ours (cell is ours
    // alone in this batch)
    if (old != cell && (old >> 32) == (cell >> 32)) P.votes[cellix] = cell;
    rep.x = 0; rep.w = rec.y;                                    // Phase2b(serverIndex, slot, round) (:1077-1081)
  }
  P.out[i] = rep;
}

// handleChosen -> choose (:1170-1197, :622-640).  in {_, server, slot, value}
__global__ void __launch_bounds__(256) vm_learn_chosen_kernel(VmParams P) {
  const Geometry& g = P.g;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  int4 rec = P.in[i];
  int server = rec.y, slot = rec.z;
  int local = local_slot(g, slot);
  if (local < 0) { report_error(P.st, FPX_ERR_SLOT_RANGE, i); return; }
  if ((uint32_t)server >= (uint32_t)g.per_group) { report_error(P.st, FPX_ERR_BAD_ACCEPTOR, i); return; }
  atomicMax(&P.votes[(size_t)local * g.voters + server], kCellChosen | (uint32_t)rec.w);   // ChosenEntry (:624)
  if (slot % g.per_group == server) {                                                       // phase2s.remove(slot) (:625)
    uint32_t* row = P.rows + (size_t)local * g.row_words;
    if (row[0] != kUnarmed) atomicOr(row, kDoneBit);
  }
}

}  // namespace fpx
