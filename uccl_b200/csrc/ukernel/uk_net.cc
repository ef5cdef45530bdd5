#include "uk_net.h"

#include <string.h>

#include <chrono>

#include "../coll/comm.h"
#include "../common/log.h"
#include "../common/timers.h"

namespace ub {

namespace {
constexpr uint32_t kUkMagic = 0x4b554e55u;  // "UNUK"
}

UkNetComm::UkNetComm(int rank, int nranks, std::shared_ptr<net::Engine> engine, std::vector<uint32_t> flows,
                     const UkNetConfig& cfg)
    : rank_(rank), n_(nranks), eng_(std::move(engine)), cfg_(cfg), arrived_((size_t)nranks * cfg.nlanes),
      expected_((size_t)nranks * cfg.nlanes, 0) {
  UB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks && (int)flows.size() == nranks, "ukernel net: bad rank / flow table");
  UB_CHECK(cfg.nlanes >= 1 && cfg.tile_bytes >= 16 && cfg.tile_bytes % 16 == 0, "ukernel net: bad lanes / tile size");
  for (auto& a : arrived_) a.store(0);
  peers_.assign((size_t)nranks, nullptr);
  for (int p = 0; p < nranks; ++p) {
    if (p == rank) continue;
    peers_[p] = new Peer();
    peers_[p]->flow = flows[p];
  }
  rx_ = std::thread([this] { receiver(); });
}

UkNetComm::~UkNetComm() {
  stop_.store(true);
  if (rx_.joinable()) rx_.join();
  // posted header receives still reference the Peer structs inside the engine: they are left allocated
}

char* UkNetComm::base(int buf) const { return bases_[buf]; }

// One thread serves all peers: a header receive is always posted; the payload receive is posted (straight into
// the destination the header names) once the header belongs to the op in flight.
void UkNetComm::receiver() {
  uint32_t idle = 0;
  while (!stop_.load(std::memory_order_relaxed)) {
    bool progress = false;
    for (int p = 0; p < n_; ++p) {
      Peer* s = peers_[p];
      if (!s) continue;
      size_t nb = 0;
      int err = 0;
      if (!s->hdr_req && !s->pay_req && !s->parked) {
        s->hdr_req = eng_->recv_async(s->flow, &s->hdr, sizeof(WireHdr));
        progress = true;
      }
      if (s->hdr_req && eng_->test(s->hdr_req, &nb, &err)) {
        s->hdr_req = nullptr;
        progress = true;
        if (err == 3) {  // orderly close by the peer: fine unless one of our Recv ops still waits for it
          s->closed.store(true);
          s->parked = true;
          s->hdr.op_seq = 0xffffffffu;
          continue;
        }
        if (err || nb != sizeof(WireHdr) || s->hdr.magic != kUkMagic || s->hdr.buf > 2 || (int)s->hdr.lane >= cfg_.nlanes) {
          if (!stop_.load()) rx_error_.store(err ? err : 100);
          s->parked = true;  // nothing more can be trusted on this flow
          s->hdr.op_seq = 0xffffffffu;
          continue;
        }
        s->parked = true;
      }
      if (s->parked && s->hdr.op_seq == cur_seq_.load(std::memory_order_acquire)) {
        s->parked = false;
        progress = true;
        if (s->hdr.bytes == 0) {
          arrived_[(size_t)p * cfg_.nlanes + s->hdr.lane].fetch_add(1, std::memory_order_release);
        } else if (s->hdr.off + s->hdr.bytes > ext_[(int)s->hdr.buf] || s->hdr.off + s->hdr.bytes < s->hdr.off) {
          // a count mismatch between ranks or a corrupt header must not overwrite memory outside this op's buffers
          rx_error_.store(102);
          s->parked = true;
          s->hdr.op_seq = 0xffffffffu;
        } else {
          s->pay_req = eng_->recv_async(s->flow, base((int)s->hdr.buf) + s->hdr.off, s->hdr.bytes);
        }
      } else if (s->parked && idle == 0) {
        ++stats_.parked_headers;
      }
      if (s->pay_req && eng_->test(s->pay_req, &nb, &err)) {
        s->pay_req = nullptr;
        progress = true;
        if (err || nb != s->hdr.bytes) rx_error_.store(err ? err : 101);
        else arrived_[(size_t)p * cfg_.nlanes + s->hdr.lane].fetch_add(1, std::memory_order_release);
      }
    }
    if (progress) {
      idle = 0;
    } else if (++idle > 200) {
      std::this_thread::sleep_for(std::chrono::microseconds(20));
    } else {
      std::this_thread::yield();
    }
  }
}

// Ready-list scheduler: an op fires when its deps are done (and, for a Recv, its message has arrived); a Send
// is done when both of its messages are acknowledged, i.e. its source may be overwritten.
void UkNetComm::run(const UkPlan& plan, char* in, char* out, int dtype, int op) {
  if (scratch_.size() < plan.scratch_bytes) scratch_.resize(plan.scratch_bytes);
  bases_[0] = in;
  bases_[1] = out;
  bases_[2] = scratch_.data();
  ext_[0] = ext_[1] = ext_[2] = 0;
  // what a peer may legally write: the scratch area the plan addresses and the collective's In / Out size
  // (`bytes` is the whole message for AllReduce / Broadcast and the per-peer block for the others); the ranges
  // my own ops touch are folded in as a lower bound
  const uint64_t whole = (plan.coll == UkColl::AllReduce || plan.coll == UkColl::Broadcast) ? plan.bytes
                                                                                              : plan.bytes * (uint64_t)plan.nranks;
  ext_[0] = ext_[1] = whole;
  ext_[2] = plan.scratch_bytes;
  for (const UkPlanOp& o : plan.ops) {
    if (o.kind == UkPlanOp::Recv) continue;
    for (const UkRef* r : {&o.dst, &o.src, &o.src2}) {
      if (o.kind != UkPlanOp::Reduce && r == &o.src2) continue;
      if (o.kind == UkPlanOp::Send && r == &o.dst) continue;  // lives in the peer's buffers
      const int b = (int)r->buf;
      if (b >= 0 && b <= 2) ext_[b] = std::max<uint64_t>(ext_[b], r->off + o.bytes);
    }
  }
  const uint32_t seq = cur_seq_.load(std::memory_order_relaxed) + 1;
  cur_seq_.store(seq, std::memory_order_release);  // parked headers of this op may now be served
  ++stats_.ops;
  auto local = [&](const UkRef& r) { return bases_[(int)r.buf] + r.off; };
  const size_t nops = plan.ops.size();
  enum : uint8_t { WAITING = 0, SENDING = 1, DONE = 2 };
  std::vector<uint8_t> state(nops, WAITING);
  struct Out {
    WireHdr hdr;
    net::Request *h = nullptr, *p = nullptr;
  };
  std::vector<Out> outs(nops);
  std::vector<uint64_t> want(nops, 0);  // Recv: arrival ordinal this op waits for
  {
    std::vector<uint64_t> ord = expected_;
    for (size_t i = 0; i < nops; ++i)
      if (plan.ops[i].kind == UkPlanOp::Recv) want[i] = ++ord[(size_t)plan.ops[i].peer * cfg_.nlanes + plan.ops[i].lane];
    expected_ = ord;
  }
  size_t done = 0;
  const uint64_t t0 = now_ns();
  uint32_t idle = 0;
  // the engine holds pointers into `outs` (headers) and the user buffers until a send completes: if this call
  // unwinds, fail those flows and wait until the engine has let go before the vector is destroyed
  struct Reaper {
    net::Engine* eng;
    std::vector<Out>& outs;
    ~Reaper() {
      for (Out& s : outs) {
        if (s.h) eng->cancel(s.h);
        if (s.p) eng->cancel(s.p);
      }
    }
  } reaper{eng_.get(), outs};
  while (done < nops) {
    bool progress = false;
    for (size_t i = 0; i < nops; ++i) {
      if (state[i] == DONE) continue;
      const UkPlanOp& o = plan.ops[i];
      if (state[i] == SENDING) {
        Out& s = outs[i];
        size_t nb = 0;
        int err = 0;
        if (s.h && eng_->test(s.h, &nb, &err)) {
          s.h = nullptr;
          UB_CHECK(!err, "ukernel net: send to rank %d failed (%d)", o.peer, err);
        }
        if (s.p && eng_->test(s.p, &nb, &err)) {
          s.p = nullptr;
          UB_CHECK(!err, "ukernel net: send to rank %d failed (%d)", o.peer, err);
        }
        if (!s.h && !s.p) state[i] = DONE, ++done, progress = true;
        continue;
      }
      bool ready = true;
      for (int d : o.deps)
        if (state[(size_t)d] != DONE) {
          ready = false;
          break;
        }
      if (!ready) continue;
      switch (o.kind) {
        case UkPlanOp::Copy:
          if (o.bytes) memmove(local(o.dst), local(o.src), o.bytes);
          state[i] = DONE, ++done, progress = true;
          break;
        case UkPlanOp::Reduce: {
          const void* srcs[2] = {local(o.src), local(o.src2)};
          host_reduce_n(local(o.dst), srcs, 2, o.bytes / (uint64_t)dtype_size(dtype), dtype, op, 1.0f);
          state[i] = DONE, ++done, progress = true;
          break;
        }
        case UkPlanOp::Send: {
          Out& s = outs[i];
          s.hdr = WireHdr{kUkMagic, seq, (uint32_t)o.dst.buf, (uint32_t)o.lane, o.dst.off, o.bytes};
          const uint32_t flow = peers_[(size_t)o.peer]->flow;
          s.h = eng_->send_async(flow, &s.hdr, sizeof(WireHdr));
          if (o.bytes) s.p = eng_->send_async(flow, local(o.src), o.bytes);
          ++stats_.sends;
          stats_.bytes_sent += o.bytes;
          state[i] = SENDING, progress = true;
          break;
        }
        case UkPlanOp::Recv:
          if (arrived_[(size_t)o.peer * cfg_.nlanes + o.lane].load(std::memory_order_acquire) >= want[i]) {
            ++stats_.recvs;
            state[i] = DONE, ++done, progress = true;
          }
          break;
        default: UB_THROW("ukernel net: unknown plan op %d", o.kind);
      }
    }
    if (progress) {
      idle = 0;
      continue;
    }
    UB_CHECK(rx_error_.load() == 0, "ukernel net: receive path failed (%d)", rx_error_.load());
    for (size_t i = 0; i < nops; ++i)
      if (state[i] == WAITING && plan.ops[i].kind == UkPlanOp::Recv && peers_[(size_t)plan.ops[i].peer]->closed.load() &&
          arrived_[(size_t)plan.ops[i].peer * cfg_.nlanes + plan.ops[i].lane].load() < want[i])
        UB_THROW("ukernel net: rank %d closed its flow while rank %d still expects data from it", plan.ops[i].peer, rank_);
    UB_CHECK(now_ns() - t0 < (uint64_t)cfg_.timeout_ms * 1000000ull, "ukernel net: %s stalled for %d ms (rank %d, %zu/%zu ops done)",
             plan.describe().c_str(), cfg_.timeout_ms, rank_, done, nops);
    if (++idle > 200) std::this_thread::sleep_for(std::chrono::microseconds(20));
    else std::this_thread::yield();
  }
}

void UkNetComm::all_reduce(const void* in, void* out, size_t count, int dtype, int op, UkAlgo algo) {
  UB_CHECK(dtype >= 0 && dtype < kNumDTypes, "ukernel net all_reduce: bad dtype %d", dtype);
  UB_CHECK(op == kSum || op == kProd || op == kMax || op == kMin, "ukernel net all_reduce: op %d unsupported", op);
  UkPlanParams p;
  p.nranks = n_, p.rank = rank_, p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  p.elem_size = (uint64_t)dtype_size(dtype), p.algo = algo;
  run(uk_plan_allreduce(count * p.elem_size, p), (char*)in, (char*)out, dtype, op);
}

void UkNetComm::all_to_all(const void* in, void* out, size_t count, int dtype) {
  UB_CHECK(in != out, "ukernel net all_to_all: in-place operation is not supported");
  UkPlanParams p;
  p.nranks = n_, p.rank = rank_, p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  run(uk_plan_alltoall(count * (uint64_t)dtype_size(dtype), p), (char*)in, (char*)out, dtype, kSum);
}

void UkNetComm::all_gather(const void* in, void* out, size_t count, int dtype) {
  UkPlanParams p;
  p.nranks = n_, p.rank = rank_, p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  run(uk_plan_allgather(count * (uint64_t)dtype_size(dtype), p), (char*)in, (char*)out, dtype, kSum);
}

void UkNetComm::reduce_scatter(const void* in, void* out, size_t count, int dtype, int op) {
  UB_CHECK(op == kSum || op == kProd || op == kMax || op == kMin, "ukernel net reduce_scatter: op %d unsupported", op);
  UkPlanParams p;
  p.nranks = n_, p.rank = rank_, p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  p.elem_size = (uint64_t)dtype_size(dtype);
  run(uk_plan_reduce_scatter(count * p.elem_size, p), (char*)in, (char*)out, dtype, op);
}

void UkNetComm::broadcast(const void* in, void* out, size_t count, int dtype, int root) {
  UkPlanParams p;
  p.nranks = n_, p.rank = rank_, p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  run(uk_plan_broadcast(count * (uint64_t)dtype_size(dtype), root, p), (char*)in, (char*)out, dtype, kSum);
}

void UkNetComm::barrier() {
  UkPlanParams p;
  p.nranks = n_, p.rank = rank_, p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  run(uk_plan_barrier(p), nullptr, nullptr, kU8, kSum);
}

}  // namespace ub
