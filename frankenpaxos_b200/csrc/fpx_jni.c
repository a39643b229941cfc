/* fpx_jni.c -- JNI entry points for the Scala side (frankenpaxos.gpu.Native).
 *
 * Every parameter is a primitive or the raw address of a direct ByteBuffer
 * (obtained on the JVM side with sun.nio.ch.DirectBuffer.address() / Unsafe), so
 * no JNIEnv function is ever called and the file compiles with the three local
 * typedefs below in an image that has no jni.h.  The JVM binds these symbols by
 * name (`System.loadLibrary("fpx")`); INTEGRATION.md has the Scala declarations.
 */
#include <stdint.h>

#include "../../include/fpx.h"

typedef void JNIEnv_;    /* opaque: never dereferenced */
typedef void* jclass_;
typedef int64_t jlong_;  /* JNI jlong  */
typedef int32_t jint_;   /* JNI jint   */

#define JNIEXPORT_ __attribute__((visibility("default")))
#define P(x) ((void*)(intptr_t)(x))

/* long create(int protocol, int f, int groups, int perGroup, int flexible, int numLeaders,
 *             int numReplicas, int slotCapacity, int overflowCapacity, int maxBatch, int device,
 *             int shardIndex, int shardCount, int numLeaderGroups)  -> handle, or the negative fpx_status */
JNIEXPORT_ jlong_ Java_frankenpaxos_gpu_Native_create(JNIEnv_* env, jclass_ cls, jint_ protocol, jint_ f,
                                                      jint_ groups, jint_ per_group, jint_ flexible,
                                                      jint_ num_leaders, jint_ num_replicas, jint_ slot_capacity,
                                                      jint_ overflow_capacity, jint_ max_batch, jint_ device,
                                                      jint_ shard_index, jint_ shard_count, jint_ num_leader_groups) {
  (void)env; (void)cls;
  fpx_config c;
  c.struct_size = (int32_t)sizeof(c);
  c.protocol = protocol; c.f = f; c.num_acceptor_groups = groups; c.acceptors_per_group = per_group;
  c.flexible = flexible; c.num_leaders = num_leaders; c.num_replicas = num_replicas;
  c.slot_capacity = slot_capacity; c.overflow_capacity = overflow_capacity; c.max_batch = max_batch;
  c.device = device; c.shard_index = shard_index; c.shard_count = shard_count;
  c.num_leader_groups = num_leader_groups;
  fpx_engine* e = 0;
  int st = fpx_create(&e, &c);
  return st == FPX_OK ? (jlong_)(intptr_t)e : (jlong_)st;
}

JNIEXPORT_ void Java_frankenpaxos_gpu_Native_destroy(JNIEnv_* env, jclass_ cls, jlong_ h) {
  (void)env; (void)cls;
  fpx_destroy((fpx_engine*)P(h));
}

/* int proxyLeaderArm(long h, long phase2aAddr, int n, long errIndexAddr) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_proxyLeaderArm(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in,
                                                             jint_ n, jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_proxyleader_arm((fpx_engine*)P(h), (const fpx_p2a*)P(in), n, (int64_t*)P(err_index));
}

/* int acceptorPhase2a(long h, long in, int n, long outPhase2b, long nPhase2b, long outNack, long nNack, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_acceptorPhase2a(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in,
                                                              jint_ n, jlong_ out_p2b, jlong_ n_p2b,
                                                              jlong_ out_nack, jlong_ n_nack, jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_acceptor_phase2a((fpx_engine*)P(h), (const fpx_p2a*)P(in), n, (fpx_p2b*)P(out_p2b),
                              (int32_t*)P(n_p2b), (fpx_nack*)P(out_nack), (int32_t*)P(n_nack),
                              (int64_t*)P(err_index));
}

/* int proxyLeaderPhase2b(long h, long in, int n, long outChosen, long nChosen, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_proxyLeaderPhase2b(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in,
                                                                 jint_ n, jlong_ out, jlong_ n_out,
                                                                 jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_proxyleader_phase2b((fpx_engine*)P(h), (const fpx_p2b*)P(in), n, (fpx_chosen*)P(out),
                                 (int32_t*)P(n_out), (int64_t*)P(err_index));
}

/* int replicaChosen(long h, long in, int n, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_replicaChosen(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in,
                                                            jint_ n, jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_replica_chosen((fpx_engine*)P(h), (const fpx_chosen*)P(in), n, (int64_t*)P(err_index));
}

/* int chosenWatermark(long h, long outAddr) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_chosenWatermark(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ out) {
  (void)env; (void)cls;
  return fpx_chosen_watermark((fpx_engine*)P(h), (int32_t*)P(out));
}

/* int snapshotAcceptor(long h, int group, int acceptor, long round, long maxVotedSlot, int firstSlot, int nSlots,
 *                      long voteRound, long voteValue)   -- what Acceptor.handlePhase1a needs (Acceptor.scala:171-179) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_snapshotAcceptor(JNIEnv_* env, jclass_ cls, jlong_ h, jint_ group,
                                                               jint_ acceptor, jlong_ round, jlong_ max_voted,
                                                               jint_ first_slot, jint_ n_slots, jlong_ vote_round,
                                                               jlong_ vote_value) {
  (void)env; (void)cls;
  return fpx_snapshot_acceptor((fpx_engine*)P(h), group, acceptor, (int32_t*)P(round), (int32_t*)P(max_voted),
                               first_slot, n_slots, (int32_t*)P(vote_round), (int32_t*)P(vote_value));
}

/* ---- S/mencius Phase2aNoopRange path and S/vanillamencius Skip (include/fpx.h) ---- */

/* int menciusArmRange(long h, long in, int n, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_menciusArmRange(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in, jint_ n,
                                                              jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_mencius_arm_range((fpx_engine*)P(h), (const fpx_p2a_range*)P(in), n, (int64_t*)P(err_index));
}

/* int menciusAcceptorNoopRange(long h, long in, int n, long out, long nOut, long outNack, long nNack, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_menciusAcceptorNoopRange(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in,
                                                                       jint_ n, jlong_ out, jlong_ n_out,
                                                                       jlong_ out_nack, jlong_ n_nack,
                                                                       jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_mencius_acceptor_noop_range((fpx_engine*)P(h), (const fpx_p2a_range*)P(in), n, (fpx_p2b_range*)P(out),
                                         (int32_t*)P(n_out), (fpx_nack*)P(out_nack), (int32_t*)P(n_nack),
                                         (int64_t*)P(err_index));
}

/* int menciusRangePhase2b(long h, long in, int n, long out, long nOut, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_menciusRangePhase2b(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in,
                                                                  jint_ n, jlong_ out, jlong_ n_out,
                                                                  jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_mencius_range_phase2b((fpx_engine*)P(h), (const fpx_p2b_range*)P(in), n, (fpx_chosen_range*)P(out),
                                   (int32_t*)P(n_out), (int64_t*)P(err_index));
}

/* int menciusReplicaChosenRange(long h, long in, int n, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_menciusReplicaChosenRange(JNIEnv_* env, jclass_ cls, jlong_ h,
                                                                        jlong_ in, jint_ n, jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_mencius_replica_chosen_range((fpx_engine*)P(h), (const fpx_chosen_range*)P(in), n,
                                          (int64_t*)P(err_index));
}

/* int vmSkip(long h, long in, int n, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_vmSkip(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in, jint_ n,
                                                     jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_vm_skip((fpx_engine*)P(h), (const fpx_vm_skip_rec*)P(in), n, (int64_t*)P(err_index));
}

/* ---- wire codec (include/fpx.h): batches of protobuf bytes <-> records ---- */

/* int wireDecodeInbound(long h, int inbound, long bytes, long offsets, int n, long kind, long out, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_wireDecodeInbound(JNIEnv_* env, jclass_ cls, jlong_ h, jint_ inbound,
                                                                jlong_ bytes, jlong_ offsets, jint_ n, jlong_ kind,
                                                                jlong_ out, jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_wire_decode_inbound((fpx_engine*)P(h), inbound, (const uint8_t*)P(bytes), (const int32_t*)P(offsets), n,
                                 (int32_t*)P(kind), (fpx_wire_rec*)P(out), (int64_t*)P(err_index));
}

/* int wireEncodePhase2b(long h, long in, int n, long out, int outCapacity, long offsets, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_wireEncodePhase2b(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in,
                                                                jint_ n, jlong_ out, jint_ out_capacity,
                                                                jlong_ offsets, jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_wire_encode_phase2b((fpx_engine*)P(h), (const fpx_p2b*)P(in), n, (uint8_t*)P(out), out_capacity,
                                 (int32_t*)P(offsets), (int64_t*)P(err_index));
}

/* int wireEncodeNack(long h, long in, int n, long out, int outCapacity, long offsets, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_wireEncodeNack(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in, jint_ n,
                                                             jlong_ out, jint_ out_capacity, jlong_ offsets,
                                                             jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_wire_encode_nack((fpx_engine*)P(h), (const fpx_nack*)P(in), n, (uint8_t*)P(out), out_capacity,
                              (int32_t*)P(offsets), (int64_t*)P(err_index));
}

/* int wireEncodeChosen(long h, long in, int n, long arena, long valueOffsets, int numValues, long out,
 *                      int outCapacity, long offsets, long err) */
JNIEXPORT_ jint_ Java_frankenpaxos_gpu_Native_wireEncodeChosen(JNIEnv_* env, jclass_ cls, jlong_ h, jlong_ in, jint_ n,
                                                               jlong_ arena, jlong_ value_offsets, jint_ num_values,
                                                               jlong_ out, jint_ out_capacity, jlong_ offsets,
                                                               jlong_ err_index) {
  (void)env; (void)cls;
  return fpx_wire_encode_chosen((fpx_engine*)P(h), (const fpx_chosen*)P(in), n, (const uint8_t*)P(arena),
                                (const int32_t*)P(value_offsets), num_values, (uint8_t*)P(out), out_capacity,
                                (int32_t*)P(offsets), (int64_t*)P(err_index));
}
