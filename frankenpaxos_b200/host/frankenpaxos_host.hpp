// frankenpaxos_host.hpp -- C++ mirror of the reference's plugin surface for the
// quorum-vote path, sitting directly on the C ABI of include/fpx.h.
//
// The reference's host side is Scala on the JVM; this image has no JVM, so the
// host layer above the C ABI is restated in C++ with the reference's names,
// argument meaning and error behaviour:
//
//   frankenpaxos::Logger            S/Logger.scala:35-118      fatal() never returns
//   frankenpaxos::FakeLogger        S/FakeLogger.scala:8-31    fatal() throws (AssertionError there)
//   frankenpaxos::Transport         S/Transport.scala:44-99    register/send/sendNoFlush/flush
//   frankenpaxos::Actor             S/Actor.scala:7-51         ctor self-registers, receive(src, bytes)
//   frankenpaxos::Chan<Dst>         S/Chan.scala:3-17          serialise with Dst's wire format, send
//   frankenpaxos::FakeTransport     S/FakeTransport.scala:64-183  message bag + deliverMessage
//   multipaxos::Config              S/multipaxos/Config.scala:6-148 (checkValid)
//   multipaxos::GpuAcceptor         drop-in for S/multipaxos/Acceptor.scala:59-220 (Phase2a path)
//   multipaxos::GpuProxyLeader      drop-in for S/multipaxos/ProxyLeader.scala:67-258
//
// Messages travel as the reference's protobuf wire bytes (proto2, scalapb):
// ProxyLeaderInbound{phase2a=1, phase2b=2}, AcceptorInbound{phase2a=2},
// ReplicaInbound{chosen=1}, LeaderInbound{nack=6}  (S/multipaxos/MultiPaxos.proto:
// 273-290, 292-298, 455-460, 525-575), so a recorded FakeTransportMessage could be
// fed in unchanged.  The actors buffer what the transport delivers and hand each
// batch, in delivery order, to a Backend: the C ABI of libfpx.so (GpuBackend), or
// -- in tests only -- the CPU oracle behind the same interface.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <algorithm>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/fpx.h"

namespace frankenpaxos {

using Bytes = std::vector<uint8_t>;
using Address = std::string;  // FakeTransportAddress(address: String), S/FakeTransport.scala:12

// ---------------------------------------------------------------- Logger
struct FatalError : std::logic_error { using std::logic_error::logic_error; };
class Logger {
 public:
  virtual ~Logger() {}
  [[noreturn]] virtual void fatal(const std::string& msg) = 0;  // S/Logger.scala:43: fatal(): Nothing
  virtual void debug(const std::string&) {}
  void check(bool b, const std::string& what = "check failed") { if (!b) fatal(what); }  // :77-80
};
class FakeLogger : public Logger {  // S/FakeLogger.scala:11-14: fatal throws
 public:
  [[noreturn]] void fatal(const std::string& msg) override { throw FatalError(msg); }
};

// ---------------------------------------------------------------- protobuf wire helpers (proto2)
namespace wire {
inline void put_varint(Bytes& b, uint64_t v) {
  while (v >= 0x80) { b.push_back((uint8_t)(v | 0x80)); v >>= 7; }
  b.push_back((uint8_t)v);
}
inline void put_int32(Bytes& b, int field, int32_t v) {  // negative int32 -> 10-byte varint
  put_varint(b, (uint64_t)(field << 3));
  put_varint(b, (uint64_t)(int64_t)v);
}
inline void put_bytes(Bytes& b, int field, const Bytes& payload) {
  put_varint(b, (uint64_t)((field << 3) | 2));
  put_varint(b, payload.size());
  b.insert(b.end(), payload.begin(), payload.end());
}
struct Reader {
  const uint8_t* p; const uint8_t* end;
  explicit Reader(const Bytes& b) : p(b.data()), end(b.data() + b.size()) {}
  Reader(const uint8_t* a, const uint8_t* z) : p(a), end(z) {}
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0; int shift = 0;
    while (p < end) { uint8_t c = *p++; v |= (uint64_t)(c & 0x7f) << shift; if (!(c & 0x80)) return v; shift += 7; }
    throw std::runtime_error("truncated varint");
  }
  // returns field number, sets wire type
  int tag(int* wt) { uint64_t t = varint(); *wt = (int)(t & 7); return (int)(t >> 3); }
  Bytes bytes() { uint64_t n = varint(); if (p + n > end) throw std::runtime_error("truncated"); Bytes b(p, p + n); p += n; return b; }
};
}  // namespace wire

// ---------------------------------------------------------------- messages on the path
struct Phase2a { int32_t slot = 0, round = 0; Bytes commandBatchOrNoop; };     // MultiPaxos.proto:273-280
struct Phase2b { int32_t groupIndex = 0, acceptorIndex = 0, slot = 0, round = 0; };  // :282-290
struct Chosen { int32_t slot = 0; Bytes commandBatchOrNoop; };                 // :292-298
struct Nack { int32_t round = 0; };                                            // :455-460

inline Bytes encode(const Phase2a& m) { Bytes b; wire::put_int32(b, 1, m.slot); wire::put_int32(b, 2, m.round); wire::put_bytes(b, 3, m.commandBatchOrNoop); return b; }
inline Bytes encode(const Phase2b& m) { Bytes b; wire::put_int32(b, 1, m.groupIndex); wire::put_int32(b, 2, m.acceptorIndex); wire::put_int32(b, 3, m.slot); wire::put_int32(b, 4, m.round); return b; }
inline Bytes encode(const Chosen& m) { Bytes b; wire::put_int32(b, 1, m.slot); wire::put_bytes(b, 2, m.commandBatchOrNoop); return b; }
inline Bytes encode(const Nack& m) { Bytes b; wire::put_int32(b, 1, m.round); return b; }
inline Bytes wrap(int field, const Bytes& inner) { Bytes b; wire::put_bytes(b, field, inner); return b; }

inline Phase2a decode_phase2a(const Bytes& b) {
  Phase2a m; wire::Reader r(b); int wt;
  while (!r.done()) { int f = r.tag(&wt); if (f == 1) m.slot = (int32_t)r.varint(); else if (f == 2) m.round = (int32_t)r.varint(); else if (f == 3) m.commandBatchOrNoop = r.bytes(); else throw std::runtime_error("Phase2a: unknown field"); }
  return m;
}
inline Phase2b decode_phase2b(const Bytes& b) {
  Phase2b m; wire::Reader r(b); int wt;
  while (!r.done()) { int f = r.tag(&wt); int32_t v = (int32_t)r.varint(); if (f == 1) m.groupIndex = v; else if (f == 2) m.acceptorIndex = v; else if (f == 3) m.slot = v; else if (f == 4) m.round = v; }
  return m;
}
inline Chosen decode_chosen(const Bytes& b) {
  Chosen m; wire::Reader r(b); int wt;
  while (!r.done()) { int f = r.tag(&wt); if (f == 1) m.slot = (int32_t)r.varint(); else m.commandBatchOrNoop = r.bytes(); }
  return m;
}
inline Nack decode_nack(const Bytes& b) { Nack m; wire::Reader r(b); int wt; while (!r.done()) { r.tag(&wt); m.round = (int32_t)r.varint(); } return m; }

// XInbound oneof wrappers: field numbers from MultiPaxos.proto:525-575
enum { kProxyLeaderPhase2a = 1, kProxyLeaderPhase2b = 2, kAcceptorPhase2a = 2, kReplicaChosen = 1, kLeaderNack = 6 };
struct Inbound { int field; Bytes body; };
inline Inbound decode_inbound(const Bytes& b) {
  wire::Reader r(b); int wt; int f = r.tag(&wt);
  if (wt != 2) throw std::runtime_error("Inbound: expected a length-delimited oneof member");
  return Inbound{f, r.bytes()};
}

// ---------------------------------------------------------------- Transport / Actor / Chan
class Actor;
class Transport {  // S/Transport.scala:44-99.  Contract: receive() calls are serial on one thread (:37-39)
 public:
  virtual ~Transport() {}
  virtual void registerActor(const Address& address, Actor* actor) = 0;             // register (:56)
  virtual void send(Actor* src, const Address& dst, const Bytes& bytes) = 0;        // (:63)
  virtual void sendNoFlush(Actor* src, const Address& dst, const Bytes& bytes) { send(src, dst, bytes); }  // (:70)
  virtual void flush(Actor*, const Address&) {}                                     // (:76)
};

class Actor {  // S/Actor.scala:7-51
 public:
  Actor(Address address, Transport& transport, Logger& logger)
      : address_(std::move(address)), transport_(transport), logger_(logger) { transport_.registerActor(address_, this); }  // :19-20
  virtual ~Actor() {}
  virtual void receive(const Address& src, const Bytes& inbound) = 0;                // :16
  const Address& address() const { return address_; }
  void send(const Address& dst, const Bytes& bytes) { transport_.send(this, dst, bytes); }  // :34
 protected:
  Address address_;
  Transport& transport_;
  Logger& logger_;
};

// Chan[Dst]: typed by the destination's Inbound wrapper (S/Chan.scala:3-17)
class Chan {
 public:
  Chan(Actor& src, Address dst) : src_(src), dst_(std::move(dst)) {}
  void send(int oneof_field, const Bytes& message) { src_.send(dst_, wrap(oneof_field, message)); }  // :12-13
  const Address& dst() const { return dst_; }
 private:
  Actor& src_;
  Address dst_;
};

// FakeTransport: the in-memory message bag (S/FakeTransport.scala:64-183)
struct FakeTransportMessage { Address src, dst; Bytes bytes; };  // :58-62
class FakeTransport : public Transport {
 public:
  explicit FakeTransport(Logger& logger) : logger_(logger) {}
  void registerActor(const Address& address, Actor* actor) override {
    if (actors.count(address)) logger_.fatal("Attempting to register an actor with address " + address + " which is already registered.");  // :80-85
    actors[address] = actor;
  }
  void send(Actor* src, const Address& dst, const Bytes& bytes) override { messages.push_back({src->address(), dst, bytes}); }  // :89-95
  // deliverMessage (:142-159): remove from the bag, hand to the destination's receive
  void deliverMessage(size_t index) {
    FakeTransportMessage m = messages.at(index);
    messages.erase(messages.begin() + (long)index);
    auto it = actors.find(m.dst);
    if (it == actors.end()) return;  // :150 a message to an unregistered actor is dropped
    it->second->receive(m.src, m.bytes);
  }
  std::map<Address, Actor*> actors;
  std::vector<FakeTransportMessage> messages;  // :74
 private:
  Logger& logger_;
};

namespace multipaxos {

// S/multipaxos/Config.scala:6-31 (address lists) + checkValid (:32-147), the clauses the path reads
struct Config {
  int f = 1;
  std::vector<Address> leaderAddresses, proxyLeaderAddresses, replicaAddresses;
  std::vector<std::vector<Address>> acceptorAddresses;
  bool flexible = false;
  int numLeaders() const { return (int)leaderAddresses.size(); }
  int numAcceptorGroups() const { return (int)acceptorAddresses.size(); }
  void checkValid() const {
    auto require = [](bool b, const char* m) { if (!b) throw std::invalid_argument(m); };  // Scala `require`
    require(f >= 1, "f must be >= 1");
    require(numLeaders() >= f + 1, "numLeaders must be >= f + 1");
    require((int)proxyLeaderAddresses.size() >= f + 1, "numProxyLeaders must be >= f + 1");
    require(numAcceptorGroups() >= 1, "numAcceptorGroups must be >= 1");
    if (!flexible) {
      for (auto& g : acceptorAddresses) require((int)g.size() == 2 * f + 1, "acceptorCluster.size must be 2*f + 1");
    } else {
      for (auto& g : acceptorAddresses) require(g.size() == acceptorAddresses[0].size(), "All row sizes must be the same");
      int n = numAcceptorGroups(), m = (int)acceptorAddresses[0].size();
      require(std::min(n, m) - 1 >= f, "An n x m grid can tolerate min(n, m) - 1 failures");
    }
    require((int)replicaAddresses.size() >= f + 1, "numReplicas must be >= f + 1");
  }
};

// What the batching actors call.  GpuBackend = the C ABI of libfpx.so.
class Backend {
 public:
  virtual ~Backend() {}
  virtual int arm(const fpx_p2a* in, int n, int64_t* err) = 0;
  virtual int phase2a(const fpx_p2a* in, int n, fpx_p2b* out, int* n_out, fpx_nack* nack, int* n_nack, int64_t* err) = 0;
  virtual int phase2b(const fpx_p2b* in, int n, fpx_chosen* out, int* n_out, int64_t* err) = 0;
  // Acceptor state read-back (Phase1b's `states.iteratorFrom`, Acceptor.scala:171-179)
  virtual void snapshot(int group, int acceptor, int* round, int* max_voted_slot, int first_slot, int n_slots,
                        int32_t* vote_round, int32_t* vote_value) = 0;
  // Phase 1 reads on the same vote cells (control path; batch boundaries by SURVEY 8(g) rule 2):
  // Acceptor.handlePhase1a (Acceptor.scala:148-182) -> -1 = round adopted, else the Nack's round;
  // Leader.safeValue over a set of responders (Leader.scala:318-329), bit g * per_group + a of `responders`
  virtual int phase1a(int group, int acceptor, int round) = 0;
  virtual void safe_values(uint32_t responders, int first_slot, int n_slots, int32_t* vote_round, int32_t* value_id,
                           int* max_slot) = 0;
  // Optional: the backend parses / serialises the hot messages itself (fpx_wire_*), so the actors
  // hand it the transport's raw bytes for Phase2b and get reply bytes back.
  virtual bool has_wire() const { return false; }
  virtual int wire_decode_proxyleader(const uint8_t*, const int32_t*, int, int32_t*, fpx_wire_rec*, int64_t*) {
    return FPX_ERR_UNSUPPORTED;
  }
  virtual int wire_encode_phase2b(const fpx_p2b*, int, uint8_t*, int, int32_t*, int64_t*) { return FPX_ERR_UNSUPPORTED; }
};
class GpuBackend : public Backend {
 public:
  GpuBackend(const Config& c, int slot_capacity, int max_batch, int device = 0) {
    c.checkValid();
    fpx_config fc;
    std::memset(&fc, 0, sizeof(fc));
    fc.struct_size = (int32_t)sizeof(fc); fc.protocol = FPX_MULTIPAXOS; fc.f = c.f;
    fc.num_acceptor_groups = c.numAcceptorGroups(); fc.acceptors_per_group = (int)c.acceptorAddresses[0].size();
    fc.flexible = c.flexible; fc.num_leaders = c.numLeaders(); fc.num_replicas = (int)c.replicaAddresses.size();
    fc.slot_capacity = slot_capacity; fc.overflow_capacity = 1024; fc.max_batch = max_batch; fc.device = device;
    fc.shard_index = 0; fc.shard_count = 1;
    int st = fpx_create(&e_, &fc);
    if (st != FPX_OK) throw std::runtime_error(std::string("fpx_create: ") + fpx_strerror(st));  // no CPU fallback
  }
  ~GpuBackend() override { fpx_destroy(e_); }
  int arm(const fpx_p2a* in, int n, int64_t* err) override { return fpx_proxyleader_arm(e_, in, n, err); }
  int phase2a(const fpx_p2a* in, int n, fpx_p2b* out, int* n_out, fpx_nack* nack, int* n_nack, int64_t* err) override {
    return fpx_acceptor_phase2a(e_, in, n, out, n_out, nack, n_nack, err);
  }
  int phase2b(const fpx_p2b* in, int n, fpx_chosen* out, int* n_out, int64_t* err) override {
    return fpx_proxyleader_phase2b(e_, in, n, out, n_out, err);
  }
  void snapshot(int group, int acceptor, int* round, int* max_voted_slot, int first_slot, int n_slots,
                int32_t* vote_round, int32_t* vote_value) override {
    fpx_snapshot_acceptor(e_, group, acceptor, round, max_voted_slot, first_slot, n_slots, vote_round, vote_value);
  }
  int phase1a(int group, int acceptor, int round) override {
    int32_t nack = -1;
    int st = fpx_acceptor_phase1a(e_, group, acceptor, round, &nack);
    if (st != FPX_OK) throw std::runtime_error(std::string("fpx_acceptor_phase1a: ") + fpx_strerror(st));
    return nack;
  }
  void safe_values(uint32_t responders, int first_slot, int n_slots, int32_t* vote_round, int32_t* value_id,
                   int* max_slot) override {
    int32_t mx = -1;
    int st = fpx_leader_safe_values(e_, responders, first_slot, n_slots, vote_round, value_id, &mx);
    if (st != FPX_OK) throw std::runtime_error(std::string("fpx_leader_safe_values: ") + fpx_strerror(st));
    *max_slot = mx;
  }
  int reset() { return fpx_reset(e_); }
  bool has_wire() const override { return true; }
  int wire_decode_proxyleader(const uint8_t* bytes, const int32_t* offs, int n, int32_t* kind, fpx_wire_rec* out,
                              int64_t* err) override {
    return fpx_wire_decode_inbound(e_, FPX_WIRE_PROXYLEADER_INBOUND, bytes, offs, n, kind, out, err);
  }
  int wire_encode_phase2b(const fpx_p2b* in, int n, uint8_t* out, int cap, int32_t* offs, int64_t* err) override {
    return fpx_wire_encode_phase2b(e_, in, n, out, cap, offs, err);
  }
  fpx_engine* engine() { return e_; }
 private:
  fpx_engine* e_ = nullptr;
};

// value_id <-> CommandBatchOrNoop bytes.  Only the id crosses the C ABI; the bytes
// stay with the host actors.
class ValueStore {
 public:
  int32_t intern(const Bytes& v) {
    auto it = ids_.find(v);
    if (it != ids_.end()) return it->second;
    int32_t id = (int32_t)values_.size();
    values_.push_back(v); ids_[v] = id;
    return id;
  }
  const Bytes& get(int32_t id) const { return values_.at((size_t)id); }
 private:
  std::vector<Bytes> values_;
  std::map<Bytes, int32_t> ids_;
};

// Drop-in for multipaxos.Acceptor's Phase2a path (S/multipaxos/Acceptor.scala:184-220).  All
// acceptors of a config that share one Backend also share one delivery buffer, so
// the engine sees their interleaved delivery stream (fpx_acceptor_phase2a's contract).
class AcceptorBatch {
 public:
  AcceptorBatch(Backend& b, ValueStore& v) : backend(b), values(v) {}
  struct Pending { fpx_p2a rec; Actor* acceptor; Address src; };
  std::vector<Pending> pending;
  Backend& backend;
  ValueStore& values;
};
class GpuAcceptor : public Actor {
 public:
  GpuAcceptor(const Address& address, Transport& transport, Logger& logger, const Config& config,
              std::shared_ptr<AcceptorBatch> batch)
      : Actor(address, transport, logger), config_(config), batch_(std::move(batch)) {
    config_.checkValid();
    for (size_t g = 0; g < config_.acceptorAddresses.size(); ++g)       // groupIndex / index (:88-90)
      for (size_t i = 0; i < config_.acceptorAddresses[g].size(); ++i)
        if (config_.acceptorAddresses[g][i] == address) { groupIndex_ = (int)g; index_ = (int)i; }
  }
  void receive(const Address& src, const Bytes& inbound) override {
    Inbound in = decode_inbound(inbound);
    if (in.field != kAcceptorPhase2a) logger_.fatal("GpuAcceptor: only Phase2a is on the accelerated path");
    Phase2a p = decode_phase2a(in.body);
    fpx_p2a rec{p.slot, p.round, batch_->values.intern(p.commandBatchOrNoop), (groupIndex_ << 16) | index_};
    batch_->pending.push_back({rec, this, src});
  }
  // One flush per delivery burst, by any acceptor sharing the batch.
  static void flush(AcceptorBatch& b, const Config& config, Logger& logger) {
    size_t n = b.pending.size();
    if (!n) return;
    std::vector<fpx_p2a> in(n);
    for (size_t i = 0; i < n; ++i) in[i] = b.pending[i].rec;
    std::vector<fpx_p2b> out(n); std::vector<fpx_nack> nack(n);
    int n_out = 0, n_nack = 0; int64_t err = -1;
    int st = b.backend.phase2a(in.data(), (int)n, out.data(), &n_out, nack.data(), &n_nack, &err);
    if (st != FPX_OK) logger.fatal(std::string("fpx status ") + fpx_strerror(st) + " at record " + std::to_string(err));
    // Replies come back compacted, each stream in delivery order: walk the batch and
    // hand every message its reply (a Phase2b carries its round, a Nack does not
    // carry the slot, so Nacks are matched by position among the rejected ones).
    // With a wire-capable backend the accepted replies are serialised in one batch
    // (ProxyLeaderInbound{phase2b} bytes, identical to wrap(kProxyLeaderPhase2b, encode(m))).
    std::vector<uint8_t> reply_bytes;
    std::vector<int32_t> reply_offs;
    if (b.backend.has_wire() && n_out > 0) {
      reply_bytes.resize((size_t)n_out * 46 + 16);
      reply_offs.resize((size_t)n_out + 1);
      int64_t werr = -1;
      int wst = b.backend.wire_encode_phase2b(out.data(), n_out, reply_bytes.data(), (int)reply_bytes.size(),
                                              reply_offs.data(), &werr);
      if (wst != FPX_OK) logger.fatal(std::string("fpx_wire_encode_phase2b: ") + fpx_strerror(wst));
    }
    size_t ip = 0, in_ = 0;
    for (size_t i = 0; i < n; ++i) {
      auto& pd = b.pending[i];
      // An acceptor's round only grows, so two identical messages (same acceptor, slot,
      // round) get the same decision unless the first is accepted: the greedy match of
      // the next unmatched Phase2b against the message is exact.
      bool accepted = ip < (size_t)n_out && out[ip].slot == pd.rec.slot && out[ip].round == pd.rec.round &&
                      ((out[ip].group << 16) | out[ip].acceptor) == pd.rec.dst;
      if (accepted) {
        if (!reply_offs.empty()) {
          pd.acceptor->send(pd.src, Bytes(reply_bytes.begin() + reply_offs[ip], reply_bytes.begin() + reply_offs[ip + 1]));
        } else {
          Phase2b m{out[ip].group, out[ip].acceptor, out[ip].slot, out[ip].round};
          pd.acceptor->send(pd.src, wrap(kProxyLeaderPhase2b, encode(m)));                   // :211-219
        }
        ++ip;
      } else {
        const fpx_nack& k = nack.at(in_++);
        pd.acceptor->send(config.leaderAddresses.at((size_t)k.leader), wrap(kLeaderNack, encode(Nack{k.round})));  // :197-198
      }
    }
    b.pending.clear();
  }
 private:
  Config config_;
  std::shared_ptr<AcceptorBatch> batch_;
  int groupIndex_ = -1, index_ = -1;
};

// Drop-in for multipaxos.ProxyLeader (S/multipaxos/ProxyLeader.scala:67-258).  Recipient
// choice (:186-197) uses the JVM's global RNG in the reference; here it is a
// caller-supplied function so that a test can fix the trace.
class GpuProxyLeader : public Actor {
 public:
  using QuorumChooser = std::function<std::vector<std::pair<int, int>>(const Phase2a&)>;  // (group, index) list
  GpuProxyLeader(const Address& address, Transport& transport, Logger& logger, const Config& config, Backend& backend,
                 ValueStore& values, QuorumChooser chooser)
      : Actor(address, transport, logger), config_(config), backend_(backend), values_(values), chooser_(std::move(chooser)) {
    config_.checkValid();
  }
  void receive(const Address& src, const Bytes& inbound) override {
    (void)src;
    Inbound in = decode_inbound(inbound);
    if (in.field == kProxyLeaderPhase2a) {                      // handlePhase2a (:175-215)
      Phase2a p = decode_phase2a(in.body);
      arms_.push_back(fpx_p2a{p.slot, p.round, values_.intern(p.commandBatchOrNoop), -1});
      // The reference forwards only the FIRST Phase2a of a (slot, round) (:177-183);
      // duplicates are dropped here the same way.
      if (seen_.insert({p.slot, p.round}).second)
        for (auto& ga : chooser_(p))
          send(config_.acceptorAddresses.at((size_t)ga.first).at((size_t)ga.second), wrap(kAcceptorPhase2a, encode(p)));
    } else if (in.field == kProxyLeaderPhase2b) {               // handlePhase2b (:217-258)
      if (backend_.has_wire()) {                                // parsed on the GPU at flush time
        raw_votes_.insert(raw_votes_.end(), inbound.begin(), inbound.end());
        raw_offs_.push_back((int32_t)raw_votes_.size());
      } else {
        Phase2b b = decode_phase2b(in.body);
        votes_.push_back(fpx_p2b{b.groupIndex, b.acceptorIndex, b.slot, b.round});
      }
    } else {
      logger_.fatal("Empty ProxyLeaderInbound encountered.");   // :166-167
    }
  }
  // arms first, then votes: sound by SURVEY.md 8(g) rule 1
  void flush() {
    int64_t err = -1;
    if (!arms_.empty()) {
      int st = backend_.arm(arms_.data(), (int)arms_.size(), &err);
      if (st != FPX_OK) logger_.fatal(std::string("fpx status ") + fpx_strerror(st) + " at record " + std::to_string(err));
      arms_.clear();
    }
    if (raw_offs_.size() > 1) {                                  // the burst's Phase2b messages, still as wire bytes
      const int n = (int)raw_offs_.size() - 1;
      std::vector<int32_t> kind((size_t)n);
      std::vector<fpx_wire_rec> rec((size_t)n);
      int st = backend_.wire_decode_proxyleader(raw_votes_.data(), raw_offs_.data(), n, kind.data(), rec.data(), &err);
      if (st != FPX_OK) logger_.fatal(std::string("fpx status ") + fpx_strerror(st) + " at message " + std::to_string(err));
      for (int i = 0; i < n; ++i) votes_.push_back(fpx_p2b{rec[(size_t)i].a, rec[(size_t)i].b, rec[(size_t)i].c, rec[(size_t)i].d});
      raw_votes_.clear();
      raw_offs_.assign(1, 0);
    }
    if (!votes_.empty()) {
      std::vector<fpx_chosen> out(votes_.size());
      int n_out = 0;
      int st = backend_.phase2b(votes_.data(), (int)votes_.size(), out.data(), &n_out, &err);
      if (st != FPX_OK)   // == the reference's logger.fatal (:220-225) / require (Grid.scala:44-47)
        logger_.fatal(std::string("fpx status ") + fpx_strerror(st) + " at record " + std::to_string(err));
      for (int i = 0; i < n_out; ++i) {                         // replicas.foreach(_.send(Chosen)) in config order (:246-253)
        Chosen c{out[i].slot, values_.get(out[i].value_id)};
        for (auto& r : config_.replicaAddresses) send(r, wrap(kReplicaChosen, encode(c)));
      }
      votes_.clear();
    }
  }
 private:
  Config config_;
  Backend& backend_;
  ValueStore& values_;
  QuorumChooser chooser_;
  std::vector<fpx_p2a> arms_;
  std::vector<fpx_p2b> votes_;
  std::vector<uint8_t> raw_votes_;                   // wire-capable backend: undecoded ProxyLeaderInbound{phase2b}
  std::vector<int32_t> raw_offs_{0};
  std::set<std::pair<int, int>> seen_;
};

}  // namespace multipaxos
}  // namespace frankenpaxos
