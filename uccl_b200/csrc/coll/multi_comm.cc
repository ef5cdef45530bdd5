// Hierarchical (NVLink inside a box, datagram rails between boxes) collectives in C++; see multi_comm.h.
#include "multi_comm.h"

#include <arpa/inet.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <thread>

#include "../common/log.h"
#include "../common/param.h"
#include "../fabric/cu_api.h"

namespace ub {

namespace {
bool float_dtype(int dt) { return dt == kF16 || dt == kF32 || dt == kF64 || dt == kBF16 || dt == kF8E4M3 || dt == kF8E5M2; }
struct RailAddr {
  uint32_t ip_be;
  uint16_t port;
  uint16_t pad;
  uint32_t listen_id;
};
inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }
// The staging copies and pinned allocations below are issued directly (not through Comm, which guards itself):
// make the communicator's GPU current for the duration of a call, whatever device the calling thread had selected.
struct DevScope {
  int prev = -1;
  bool active = false;
  explicit DevScope(const MultiComm& m) {
    if (m.is_host()) return;
    cudaGetDevice(&prev);
    if (prev != m.device()) {
      cudaSetDevice(m.device());
      active = true;
    }
  }
  ~DevScope() {
    if (active) cudaSetDevice(prev);
  }
};
}  // namespace

std::shared_ptr<MultiComm> MultiComm::create(const UniqueId& id, int rank, int nranks, int local_size, int device,
                                             const CommConfig& cfg) {
  UB_CHECK(local_size >= 1 && local_size <= kMaxRanks && nranks % local_size == 0,
           "multi-box communicator: %d ranks are not a multiple of the box size %d", nranks, local_size);
  auto m = std::shared_ptr<MultiComm>(new MultiComm());
  m->rank_ = rank;
  m->nranks_ = nranks;
  m->L_ = local_size;
  m->N_ = nranks / local_size;
  m->node_ = rank / local_size;
  m->lrank_ = rank % local_size;
  m->timeout_ms_ = (int)param_load("MN_TIMEOUT_MS", 120000);
  m->pipeline_bytes_ = (size_t)param_load("MN_PIPELINE_BYTES", 8 << 20);
  m->small_bytes_ = (size_t)param_load("MN_SMALL_BYTES", 16 << 10);
  m->scratch_cap_ = (size_t)param_load("MN_SCRATCH_MB", 32) << 20;
  Bootstrap g(id, rank, nranks);
  // one rendezvous per box for its NVLink communicator (the relay lives on the box: loopback)
  UniqueId mine;
  memset(&mine, 0, sizeof(mine));
  if (m->lrank_ == 0) mine = Bootstrap::create_id("127.0.0.1");
  std::vector<UniqueId> ids((size_t)nranks);
  g.allgather(&mine, ids.data(), sizeof(UniqueId));
  m->local_ = Comm::create(ids[(size_t)m->node_ * local_size], m->lrank_, local_size, device, cfg);
  // one rail per local rank
  // pick the rail's NIC FIRST and bind the engine to it: an engine on 0.0.0.0 would leave through the routing
  // table's default interface with a source address other than the advertised one (peers drop such packets),
  // and would skip the MTU clamp / NIC-local CPU pinning.  UCCL_B200_NET_BIND_IP still overrides.
  net::EngineConfig ec = net::EngineConfig::from_env();
  std::string ip = ec.bind_ip;
  if (ip.empty() || ip == "0.0.0.0") {
    auto ifs = net::list_interfaces();
    ip = ifs.empty() ? "127.0.0.1" : ifs[(size_t)m->lrank_ % ifs.size()].second;
    ec.bind_ip = ip;
  }
  m->engine_.reset(new net::Engine(ec));
  RailAddr me{};
  inet_pton(AF_INET, ip.c_str(), &me.ip_be);
  me.port = m->engine_->port();
  me.listen_id = m->engine_->listen();
  std::vector<RailAddr> addrs((size_t)nranks);
  g.allgather(&me, addrs.data(), sizeof(RailAddr));
  m->rail_.assign((size_t)m->N_, 0);
  for (int k = 0; k < m->node_; ++k) {  // higher node connects to lower node of the same rail
    const RailAddr& a = addrs[(size_t)k * local_size + m->lrank_];
    char ipb[INET_ADDRSTRLEN];
    inet_ntop(AF_INET, &a.ip_be, ipb, sizeof(ipb));
    const uint32_t f = m->engine_->connect(ipb, a.port, a.listen_id, m->timeout_ms_);
    uint32_t who = (uint32_t)m->node_;
    m->wait_req(m->engine_->send_async(f, &who, sizeof(who)), "rail hello");
    m->rail_[(size_t)k] = f;
  }
  for (int i = m->node_ + 1; i < m->N_; ++i) {
    const uint32_t f = m->engine_->accept(me.listen_id, m->timeout_ms_);
    uint32_t who = 0;
    m->wait_req(m->engine_->recv_async(f, &who, sizeof(who)), "rail hello");
    UB_CHECK(who < (uint32_t)m->N_ && (int)who != m->node_, "rail hello from unknown node %u", who);
    m->rail_[who] = f;
  }
  m->engine_->close_listen(me.listen_id);
  g.barrier();
  UB_INFO(SUB_INIT, "multi-box communicator: rank %d/%d = box %d/%d local %d/%d, rail via %s", rank, nranks, m->node_,
          m->N_, m->lrank_, m->L_, ip.c_str());
  return m;
}

MultiComm::~MultiComm() {
  for (auto& d : dbuf_)
    if (d.p && local_) local_->free(d.p);
  engine_.reset();  // graceful: FIN exchange + linger
  for (auto& h : hbuf_) {
    if (!h.p) continue;
    if (h.pinned) cudaFreeHost(h.p);
    else ::free(h.p);
  }
}

std::string MultiComm::describe() const {
  char b[256];
  snprintf(b, sizeof(b), "MultiComm rank %d/%d box %d/%d local %d/%d | ", rank_, nranks_, node_, N_, lrank_, L_);
  return std::string(b) + local_->describe();
}

// ------------------------------------------------------------------------------------------ staging
char* MultiComm::host_stage(size_t bytes, int slot) {
  HostBuf& h = hbuf_[slot];
  if (h.cap >= bytes && h.p) return h.p;
  if (h.p) {
    if (h.pinned) cudaFreeHost(h.p);
    else ::free(h.p);
  }
  const size_t cap = std::max<size_t>(bytes, 1 << 16) * 5 / 4;
  h.pinned = !is_host();
  if (h.pinned) UB_CUDA(cudaHostAlloc((void**)&h.p, cap, cudaHostAllocDefault));
  else h.p = (char*)::malloc(cap);
  UB_CHECK(h.p, "host staging allocation of %zu bytes failed", cap);
  h.cap = cap;
  return h.p;
}

void* MultiComm::scratch(size_t bytes, int slot) {
  DevBuf& d = dbuf_[slot];
  if (d.cap >= bytes && d.p) return d.p;
  if (d.p) local_->free(d.p);
  const size_t cap = std::max<size_t>(bytes, 1 << 16) * 5 / 4;  // same growth on every local rank: collective alloc
  d.p = local_->alloc(cap);
  d.cap = cap;
  return d.p;
}

void MultiComm::to_host(void* host, const void* dev, size_t bytes, cudaStream_t s) {
  if (!bytes || host == dev) return;
  if (is_host()) memcpy(host, dev, bytes);
  else UB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s));
}
void MultiComm::to_dev(void* dev, const void* host, size_t bytes, cudaStream_t s) {
  if (!bytes || host == dev) return;
  if (is_host()) memcpy(dev, host, bytes);
  else UB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s));
}
void MultiComm::copy_dd(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (!bytes || dst == src) return;
  if (is_host()) memmove(dst, src, bytes);
  else UB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
}
void MultiComm::zero(void* p, size_t bytes, cudaStream_t s) {
  if (!bytes) return;
  if (is_host()) memset(p, 0, bytes);
  else UB_CUDA(cudaMemsetAsync(p, 0, bytes, s));
}
void MultiComm::sync(cudaStream_t s) {
  if (!is_host()) UB_CUDA(cudaStreamSynchronize(s));
}

// ------------------------------------------------------------------------------------ rail primitives
void MultiComm::wait_req(net::Request* r, const char* what) {
  size_t n = 0;
  UB_CHECK(engine_->wait(r, &n, timeout_ms_), "multi-box %s failed or timed out after %d ms (box %d, rail %d)", what,
           timeout_ms_, node_, lrank_);
}

void MultiComm::rail_sendrecv(const void* sbuf, size_t sbytes, int to_node, void* rbuf, size_t rbytes, int from_node) {
  net::Request* rr = engine_->recv_async(rail_[(size_t)from_node], rbuf, rbytes);
  net::Request* sr = engine_->send_async(rail_[(size_t)to_node], sbuf, sbytes);
  wait_req(rr, "rail receive");
  wait_req(sr, "rail send");
}

void MultiComm::rail_allreduce(void* buf, size_t count, int dtype, int op) {
  const int n = N_, r = node_;
  if (n == 1 || count == 0) return;
  const size_t es = (size_t)dtype_size(dtype);
  if ((n & (n - 1)) == 0 && count * es <= small_bytes_) {  // log2(n) exchanges of the whole vector
    char* tmp = host_stage(count * es, 3);
    for (int d = 1; d < n; d <<= 1) {
      rail_sendrecv(buf, count * es, r ^ d, tmp, count * es, r ^ d);
      const void* srcs[2] = {buf, tmp};
      host_reduce_n(buf, srcs, 2, count, dtype, op, 1.0f);
    }
    return;
  }
  auto lo = [&](int k) { return (size_t)k * count / (size_t)n; };
  size_t mx = 0;
  for (int k = 0; k < n; ++k) mx = std::max(mx, lo(k + 1) - lo(k));
  char* tmp = host_stage(mx * es, 3);
  char* b = static_cast<char*>(buf);
  const int nxt = (r + 1) % n, prv = (r - 1 + n) % n;
  for (int step = 0; step < n - 1; ++step) {  // reduce-scatter: box r ends up owning segment (r+1) % n
    const int si = (r - step + n) % n, ri = (r - step - 1 + 2 * n) % n;
    const size_t sc = lo(si + 1) - lo(si), rc = lo(ri + 1) - lo(ri);
    rail_sendrecv(b + lo(si) * es, sc * es, nxt, tmp, rc * es, prv);
    const void* srcs[2] = {b + lo(ri) * es, tmp};
    host_reduce_n(b + lo(ri) * es, srcs, 2, rc, dtype, op, 1.0f);
  }
  for (int step = 0; step < n - 1; ++step) {  // all-gather of the reduced segments
    const int si = (r + 1 - step + 2 * n) % n, ri = (r - step + 2 * n) % n;
    rail_sendrecv(b + lo(si) * es, (lo(si + 1) - lo(si)) * es, nxt, b + lo(ri) * es, (lo(ri + 1) - lo(ri)) * es, prv);
  }
}

void MultiComm::rail_allgather(void* buf, size_t bytes) {
  const int n = N_, r = node_;
  char* b = static_cast<char*>(buf);
  const int nxt = (r + 1) % n, prv = (r - 1 + n) % n;
  for (int step = 0; step < n - 1; ++step) {
    const int si = (r - step + 2 * n) % n, ri = (r - step - 1 + 2 * n) % n;
    rail_sendrecv(b + (size_t)si * bytes, bytes, nxt, b + (size_t)ri * bytes, bytes, prv);
  }
}

void MultiComm::rail_reduce_scatter(void* buf, size_t count, int dtype, int op, void* out) {
  const int n = N_, r = node_;
  const size_t es = (size_t)dtype_size(dtype), bytes = count * es;
  char* b = static_cast<char*>(buf);
  char* tmp = host_stage(bytes, 3);
  const int nxt = (r + 1) % n, prv = (r - 1 + n) % n;
  for (int step = 0; step < n - 1; ++step) {  // ring shifted so that the block finished at box r is block r
    const int si = (r - step - 1 + 2 * n) % n, ri = (r - step - 2 + 2 * n) % n;
    rail_sendrecv(b + (size_t)si * bytes, bytes, nxt, tmp, bytes, prv);
    const void* srcs[2] = {b + (size_t)ri * bytes, tmp};
    host_reduce_n(b + (size_t)ri * bytes, srcs, 2, count, dtype, op, 1.0f);
  }
  memcpy(out, b + (size_t)r * bytes, bytes);
}

void MultiComm::rail_broadcast(void* buf, size_t bytes, int root_node) {
  const int n = N_;
  const int vr = (node_ - root_node + n) % n;  // binomial tree relative to the root box
  int mask = 1;
  while (mask < n) {
    if (vr & mask) {
      wait_req(engine_->recv_async(rail_[(size_t)((vr - mask + root_node) % n)], buf, bytes), "rail broadcast receive");
      break;
    }
    mask <<= 1;
  }
  mask >>= 1;
  while (mask > 0) {
    if (vr + mask < n) wait_req(engine_->send_async(rail_[(size_t)((vr + mask + root_node) % n)], buf, bytes), "rail broadcast send");
    mask >>= 1;
  }
}

void MultiComm::rail_alltoall(const void* in, void* out, size_t bytes) {
  const int n = N_, r = node_;
  const char* i = static_cast<const char*>(in);
  char* o = static_cast<char*>(out);
  memcpy(o + (size_t)r * bytes, i + (size_t)r * bytes, bytes);
  for (int d = 1; d < n; ++d) {
    const int to = (r + d) % n, from = (r - d + n) % n;
    rail_sendrecv(i + (size_t)to * bytes, bytes, to, o + (size_t)from * bytes, bytes, from);
  }
}

void MultiComm::rail_barrier() {
  const int n = N_, r = node_;
  char tok = 1, got = 0;
  for (int d = 1; d < n; d <<= 1) rail_sendrecv(&tok, 1, (r + d) % n, &got, 1, (r - d % n + n) % n);
}

// ---------------------------------------------------------------------------------------- collectives
void MultiComm::allreduce(const void* in, void* out, size_t count, int dtype, int op, cudaStream_t st, float scale) {
  DevScope dev_scope(*this);
  UB_CHECK(dtype >= 0 && dtype < kNumDTypes && op >= 0 && op < kNumOps, "allreduce: bad dtype/op");
  if (count == 0) return;
  UB_CHECK((op != kAvg && scale == 1.0f) || float_dtype(dtype), "allreduce across boxes: avg / scale need a floating-point dtype");
  if (N_ == 1) {
    ArOpts o;
    o.scale = scale;
    local_->allreduce(in, out, count, dtype, op, st, o);
    return;
  }
  const int inner = op == kAvg ? kSum : op;
  const float sc = scale * (op == kAvg ? 1.0f / (float)nranks_ : 1.0f);
  const size_t es = (size_t)dtype_size(dtype), per = ceil_div(count, (size_t)L_);
  if (count * es <= small_bytes_) {
    // latency bound: two phases instead of three -- one NVLink all-reduce, then EVERY local rank all-reduces the
    // whole (tiny) vector along its rail (L x the bytes on the network, one kernel and one staging hop less)
    char* T = static_cast<char*>(scratch(count * es, 1));
    if (L_ > 1) {
      ArOpts o;
      local_->allreduce(in, T, count, dtype, inner, st, o);
    } else {
      copy_dd(T, in, count * es, st);
    }
    char* H = is_host() ? T : host_stage(count * es, 0);
    to_host(H, T, count * es, st);
    sync(st);
    rail_allreduce(H, count, dtype, inner);
    if (sc != 1.0f) {
      const void* one[1] = {H};
      host_reduce_n(H, one, 1, count, dtype, kSum, sc);
    }
    to_dev(out, H, count * es, st);
    if (!is_host()) sync(st);
    return;
  }
  if (per * es > pipeline_bytes_) {  // large: block pipeline straight from `in` to `out` (scratch = a few blocks)
    allreduce_pipelined(static_cast<const char*>(in), static_cast<char*>(out), count, per, dtype, inner, sc, st);
    return;
  }
  char* W = static_cast<char*>(scratch(per * L_ * es, 0));
  char* S = static_cast<char*>(scratch(per * es, 1));
  copy_dd(W, in, count * es, st);
  zero(W + count * es, (per * L_ - count) * es, st);
  if (L_ > 1) local_->reduce_scatter(W, S, per, dtype, inner, st, 1.0f);
  else copy_dd(S, W, per * es, st);
  char* H = is_host() ? S : host_stage(per * es, 0);
  to_host(H, S, per * es, st);
  sync(st);
  rail_allreduce(H, per, dtype, inner);
  if (sc != 1.0f) {
    const void* one[1] = {H};
    host_reduce_n(H, one, 1, per, dtype, kSum, sc);
  }
  to_dev(S, H, per * es, st);
  if (L_ > 1) local_->allgather(S, W, per, dtype, st);
  else copy_dd(W, S, per * es, st);
  copy_dd(out, W, count * es, st);
  if (!is_host()) sync(st);  // the pinned staging buffer and the scratch are reused by the next call (maybe on another stream)
}

// Block b = columns [lo, hi) of every row of work[L][per].  While block b is on the rail (helper thread), this
// thread runs the NVLink reduce-scatter of block b+1 and the NVLink all-gather of block b-1.
void MultiComm::allreduce_pipelined(const char* in, char* out, size_t count, size_t per, int dtype, int op, float sc,
                                    cudaStream_t st) {
  const size_t es = (size_t)dtype_size(dtype);
  const size_t cols = std::max<size_t>(1, pipeline_bytes_ / es);
  const size_t nblk = ceil_div(per, cols);
  char* B[2] = {static_cast<char*>(scratch((size_t)L_ * cols * es, 6)), static_cast<char*>(scratch((size_t)L_ * cols * es, 7))};
  char* S[2] = {static_cast<char*>(scratch(cols * es, 8)), static_cast<char*>(scratch(cols * es, 9))};
  char* H[2] = {is_host() ? S[0] : host_stage(cols * es, 4), is_host() ? S[1] : host_stage(cols * es, 5)};
  (void)host_stage(ceil_div(cols, (size_t)N_) * es + 64, 3);  // the rail's receive buffer: allocate before threads run
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (!is_host())
    for (auto& e : ev) UB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  auto width = [&](size_t b) { return std::min(cols, per - b * cols); };
  auto stage_a = [&](size_t b) {  // gather the block, reduce inside the box, start the download
    const size_t w = width(b), lo = b * cols;
    const int k = (int)(b & 1);
    for (int l = 0; l < L_; ++l) {  // row l of the (virtually padded) [L][per] view of `in`
      const size_t g0 = (size_t)l * per + lo;
      const size_t valid = g0 >= count ? 0 : std::min(w, count - g0);
      copy_dd(B[k] + (size_t)l * w * es, in + g0 * es, valid * es, st);
      zero(B[k] + ((size_t)l * w + valid) * es, (w - valid) * es, st);
    }
    if (L_ > 1) local_->reduce_scatter(B[k], S[k], w, dtype, op, st, 1.0f);
    else copy_dd(S[k], B[k], w * es, st);
    to_host(H[k], S[k], w * es, st);
    if (!is_host()) UB_CUDA(cudaEventRecord(ev[k], st));
  };
  auto stage_c = [&](size_t b) {  // upload, re-assemble inside the box, scatter the block back
    const size_t w = width(b), lo = b * cols;
    const int k = (int)(b & 1);
    to_dev(S[k], H[k], w * es, st);
    if (L_ > 1) local_->allgather(S[k], B[k], w, dtype, st);
    else copy_dd(B[k], S[k], w * es, st);
    for (int l = 0; l < L_; ++l) {
      const size_t g0 = (size_t)l * per + lo;
      const size_t valid = g0 >= count ? 0 : std::min(w, count - g0);
      copy_dd(out + g0 * es, B[k] + (size_t)l * w * es, valid * es, st);
    }
  };
  std::exception_ptr net_err;
  std::thread net;
  auto start_net = [&](size_t b) {
    net = std::thread([this, b, &ev, &H, &net_err, width, dtype, op, sc] {
      try {
        const int k = (int)(b & 1);
        if (!is_host()) UB_CUDA(cudaEventSynchronize(ev[k]));
        rail_allreduce(H[k], width(b), dtype, op);
        if (sc != 1.0f) {
          const void* one[1] = {H[k]};
          host_reduce_n(H[k], one, 1, width(b), dtype, kSum, sc);
        }
      } catch (...) {
        net_err = std::current_exception();
      }
    });
  };
  auto join_net = [&] {
    if (net.joinable()) net.join();
    if (net_err) {
      if (!is_host())
        for (auto& e : ev) cudaEventDestroy(e);
      std::rethrow_exception(net_err);
    }
  };
  stage_a(0);
  start_net(0);
  for (size_t b = 1; b < nblk; ++b) {
    // block b reuses the buffers of block b-2, whose upload (stage_c) was issued on this stream before: ordered
    stage_a(b);
    join_net();      // block b-1 is off the rail
    start_net(b);
    stage_c(b - 1);  // overlaps block b's network time
  }
  join_net();
  stage_c(nblk - 1);
  if (!is_host()) {
    sync(st);  // pinned buffers and events are reused / destroyed
    for (auto& e : ev) cudaEventDestroy(e);
  }
}

// all_gather / reduce_scatter / all_to_all work on [world][count] layouts.  They run in column chunks of at most
// scratch_cap_ bytes per buffer, so the symmetric-heap scratch stays small however large the message is: chunk
// [off, off+m) of every rank block is gathered into dense scratch, exchanged, and scattered back.
void MultiComm::allgather(const void* in, void* out, size_t count, int dtype, cudaStream_t st) {
  DevScope dev_scope(*this);
  if (count == 0) return;
  if (N_ == 1) {
    local_->allgather(in, out, count, dtype, st);
    return;
  }
  const size_t es = (size_t)dtype_size(dtype), W = (size_t)nranks_;
  const size_t step = std::max<size_t>(1, scratch_cap_ / (W * es));
  const char* i = static_cast<const char*>(in);
  char* o = static_cast<char*>(out);
  for (size_t off = 0; off < count; off += step) {
    const size_t m = std::min(step, count - off), mb = m * es;
    char* H = host_stage((size_t)N_ * mb, 0);
    to_host(H + (size_t)node_ * mb, i + off * es, mb, st);
    sync(st);
    rail_allgather(H, mb);
    char* R = static_cast<char*>(scratch((size_t)N_ * mb, 0));
    to_dev(R, H, (size_t)N_ * mb, st);
    if (L_ > 1) {
      char* G = static_cast<char*>(scratch((size_t)L_ * N_ * mb, 1));
      local_->allgather(R, G, (size_t)N_ * m, dtype, st);
      for (int k = 0; k < N_; ++k)  // G is [local rank][box], the result is [box][local rank]
        for (int l = 0; l < L_; ++l)
          copy_dd(o + (((size_t)k * L_ + l) * count + off) * es, G + ((size_t)l * N_ + k) * mb, mb, st);
    } else {
      for (int k = 0; k < N_; ++k) copy_dd(o + ((size_t)k * count + off) * es, R + (size_t)k * mb, mb, st);
    }
    if (!is_host()) sync(st);  // the pinned buffer is reused by the next chunk / call
  }
}

void MultiComm::reduce_scatter(const void* in, void* out, size_t count, int dtype, int op, cudaStream_t st) {
  DevScope dev_scope(*this);
  if (count == 0) return;
  UB_CHECK(op != kAvg || float_dtype(dtype), "reduce_scatter across boxes: avg needs a floating-point dtype");
  if (N_ == 1) {
    local_->reduce_scatter(in, out, count, dtype, op, st);
    return;
  }
  const int inner = op == kAvg ? kSum : op;
  const float sc = op == kAvg ? 1.0f / (float)nranks_ : 1.0f;
  const size_t es = (size_t)dtype_size(dtype), W = (size_t)nranks_;
  const size_t step = std::max<size_t>(1, scratch_cap_ / (W * es));
  const char* i = static_cast<const char*>(in);
  char* o = static_cast<char*>(out);
  for (size_t off = 0; off < count; off += step) {
    const size_t m = std::min(step, count - off), mb = m * es;
    char* part = static_cast<char*>(scratch((size_t)N_ * mb, 1));
    if (L_ > 1) {
      char* P = static_cast<char*>(scratch((size_t)L_ * N_ * mb, 0));  // [dst local rank][box] <- in [box][local rank]
      for (int k = 0; k < N_; ++k)
        for (int l = 0; l < L_; ++l)
          copy_dd(P + ((size_t)l * N_ + k) * mb, i + (((size_t)k * L_ + l) * count + off) * es, mb, st);
      local_->reduce_scatter(P, part, (size_t)N_ * m, dtype, inner, st, 1.0f);
    } else {
      for (int k = 0; k < N_; ++k) copy_dd(part + (size_t)k * mb, i + ((size_t)k * count + off) * es, mb, st);
    }
    char* H = is_host() ? part : host_stage((size_t)N_ * mb, 0);
    to_host(H, part, (size_t)N_ * mb, st);
    sync(st);
    char* O = host_stage(mb, 1);
    rail_reduce_scatter(H, m, dtype, inner, O);
    if (sc != 1.0f) {
      const void* one[1] = {O};
      host_reduce_n(O, one, 1, m, dtype, kSum, sc);
    }
    to_dev(o + off * es, O, mb, st);
    if (!is_host()) sync(st);
  }
}

void MultiComm::broadcast(const void* in, void* out, size_t count, int dtype, int root, cudaStream_t st) {
  DevScope dev_scope(*this);
  UB_CHECK(root >= 0 && root < nranks_, "broadcast: bad root %d", root);
  if (count == 0) return;
  if (N_ == 1) {
    local_->broadcast(in, out, count, dtype, root, st);
    return;
  }
  const size_t b = count * (size_t)dtype_size(dtype);
  const int root_node = root / L_, root_l = root % L_;
  char* o = static_cast<char*>(out);
  if (node_ == root_node) {
    if (L_ > 1) local_->broadcast(in, out, count, dtype, root_l, st);
    else copy_dd(out, in, b, st);
  }
  // rail l carries bytes [l*per, (l+1)*per) of every chunk to the other boxes, which re-assemble over NVLink
  for (size_t base = 0; base < b; base += scratch_cap_ ? scratch_cap_ : b) {
    const size_t nb = std::min(scratch_cap_ ? scratch_cap_ : b, b - base);
    const size_t per = (ceil_div(nb, (size_t)L_) + 15) / 16 * 16;
    const size_t lo = std::min((size_t)lrank_ * per, nb), hi = std::min(lo + per, nb);
    char* H = host_stage(per, 0);
    if (node_ == root_node) {
      to_host(H, o + base + lo, hi - lo, st);
      sync(st);
    }
    rail_broadcast(H, per, root_node);
    if (node_ != root_node) {
      char* P = static_cast<char*>(scratch(per, 0));
      to_dev(P, H, per, st);
      if (L_ > 1) {
        char* F = static_cast<char*>(scratch(per * L_, 1));
        local_->allgather(P, F, per, kU8, st);
        copy_dd(o + base, F, nb, st);
      } else {
        copy_dd(o + base, P, nb, st);
      }
      if (!is_host()) sync(st);
    }
  }
}

void MultiComm::reduce(const void* in, void* out, size_t count, int dtype, int op, int root, cudaStream_t st) {
  DevScope dev_scope(*this);
  UB_CHECK(root >= 0 && root < nranks_, "reduce: bad root %d", root);
  if (count == 0) return;
  const size_t b = count * (size_t)dtype_size(dtype);
  void* T = scratch(b, 2);
  allreduce(in, T, count, dtype, op, st);
  if (rank_ == root) copy_dd(out, T, b, st);
}

void MultiComm::alltoall(const void* in, void* out, size_t count, int dtype, cudaStream_t st) {
  DevScope dev_scope(*this);
  if (count == 0) return;
  if (N_ == 1) {
    local_->alltoall(in, out, count, dtype, st);
    return;
  }
  const size_t es = (size_t)dtype_size(dtype), W = (size_t)nranks_;
  const size_t step = std::max<size_t>(1, scratch_cap_ / (W * es));
  const char* i = static_cast<const char*>(in);
  char* o = static_cast<char*>(out);
  for (size_t off = 0; off < count; off += step) {
    const size_t m = std::min(step, count - off), mb = m * es;
    char* C = static_cast<char*>(scratch(W * mb, 2));  // [dst box][src local]: what the rail exchanges
    if (L_ > 1) {
      char* A = static_cast<char*>(scratch(W * mb, 0));  // [dst local][dst box]
      char* B = static_cast<char*>(scratch(W * mb, 1));  // [src local][dst box] after the NVLink hop
      for (int k = 0; k < N_; ++k)
        for (int l = 0; l < L_; ++l)
          copy_dd(A + ((size_t)l * N_ + k) * mb, i + (((size_t)k * L_ + l) * count + off) * es, mb, st);
      local_->alltoall(A, B, (size_t)N_ * m, dtype, st);
      for (int l = 0; l < L_; ++l)
        for (int k = 0; k < N_; ++k) copy_dd(C + ((size_t)k * L_ + l) * mb, B + ((size_t)l * N_ + k) * mb, mb, st);
    } else {
      for (int k = 0; k < N_; ++k) copy_dd(C + (size_t)k * mb, i + ((size_t)k * count + off) * es, mb, st);
    }
    char* H = host_stage(W * mb, 0);
    to_host(H, C, W * mb, st);
    sync(st);
    char* O = host_stage(W * mb, 1);
    rail_alltoall(H, O, (size_t)L_ * mb);  // result is [src box][src local] = global source order
    for (size_t s_ = 0; s_ < W; ++s_) to_dev(o + (s_ * count + off) * es, O + s_ * mb, mb, st);
    if (!is_host()) sync(st);
  }
}

void MultiComm::alltoallv(const void* in, const size_t* sc, const size_t* sd, void* out, const size_t* rc, const size_t* rd,
                          int dtype, cudaStream_t st) {
  DevScope dev_scope(*this);
  if (N_ == 1) {
    local_->alltoallv(in, sc, sd, out, rc, rd, dtype, st);
    return;
  }
  const size_t es = (size_t)dtype_size(dtype), W = (size_t)nranks_;
  // global maximum chunk (one scalar max-all-reduce), then the padded equal-split exchange
  long long mx = 0;
  for (size_t d = 0; d < W; ++d) mx = std::max<long long>(mx, (long long)sc[d]);
  long long* dm = static_cast<long long*>(scratch(64, 3));
  to_dev(dm, &mx, sizeof(mx), st);
  sync(st);  // `mx` is a stack variable
  allreduce(dm, dm, 1, kI64, kMax, st);
  to_host(&mx, dm, sizeof(mx), st);
  sync(st);
  if (mx == 0) return;
  const size_t m = (size_t)mx;
  char* pin = static_cast<char*>(scratch(W * m * es, 4));
  char* pout = static_cast<char*>(scratch(W * m * es, 5));
  const char* i = static_cast<const char*>(in);
  char* o = static_cast<char*>(out);
  for (size_t d = 0; d < W; ++d) copy_dd(pin + d * m * es, i + sd[d] * es, sc[d] * es, st);
  alltoall(pin, pout, m, dtype, st);
  for (size_t s = 0; s < W; ++s) copy_dd(o + rd[s] * es, pout + s * m * es, rc[s] * es, st);
  if (!is_host()) sync(st);
}

void MultiComm::barrier(cudaStream_t st) {
  DevScope dev_scope(*this);
  local_->barrier(st);
  sync(st);
  rail_barrier();
  local_->barrier(st);
  sync(st);
}

void MultiComm::group_p2p(const std::vector<Comm::P2pOp>& ops, cudaStream_t st) {
  DevScope dev_scope(*this);
  std::vector<Comm::P2pOp> local_ops;
  struct NetOp {
    Comm::P2pOp op;
    int node;
    std::vector<char> stage;
    net::Request* req = nullptr;
  };
  std::vector<NetOp> net_ops;
  for (const auto& o : ops) {
    UB_CHECK(o.peer >= 0 && o.peer < nranks_, "send/recv: bad peer %d", o.peer);
    const int pn = o.peer / L_, pl = o.peer % L_;
    if (pn == node_) {
      Comm::P2pOp lo = o;
      lo.peer = pl;
      local_ops.push_back(lo);
    } else {
      UB_CHECK(pl == lrank_, "send/recv between different rails of different boxes is not routed (rank %d -> %d): go through "
               "the peer's rail-mate on this box", rank_, o.peer);
      NetOp n;
      n.op = o;
      n.node = pn;
      net_ops.push_back(std::move(n));
    }
  }
  // receives first so that data lands in place, then sends; the NVLink part runs concurrently on the stream
  for (auto& n : net_ops)
    if (!n.op.is_send) {
      void* dst = n.op.buf;
      if (!is_host()) {
        n.stage.resize(n.op.bytes);
        dst = n.stage.data();
      }
      n.req = engine_->recv_async(rail_[(size_t)n.node], dst, n.op.bytes);
    }
  bool staged = false;
  for (auto& n : net_ops)
    if (n.op.is_send && !is_host()) {
      n.stage.resize(n.op.bytes);
      to_host(n.stage.data(), n.op.buf, n.op.bytes, st);
      staged = true;
    }
  if (staged) sync(st);
  for (auto& n : net_ops)
    if (n.op.is_send) n.req = engine_->send_async(rail_[(size_t)n.node], is_host() ? n.op.buf : (void*)n.stage.data(), n.op.bytes);
  if (!local_ops.empty()) local_->group_p2p(local_ops, st);
  for (auto& n : net_ops) wait_req(n.req, n.op.is_send ? "send" : "recv");
  bool up = false;
  for (auto& n : net_ops)
    if (!n.op.is_send && !is_host()) {
      to_dev(n.op.buf, n.stage.data(), n.op.bytes, st);
      up = true;
    }
  if (up) sync(st);  // pageable staging vectors die with this call
}

}  // namespace ub
