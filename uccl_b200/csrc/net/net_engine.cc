// Engine thread of the inter-node datagram transport; see net_engine.h for the design.
#include "net_engine.h"

#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <ifaddrs.h>
#include <net/if.h>
#include <pthread.h>
#include <sched.h>
#include <string.h>
#include <sys/epoll.h>
#include <sys/ioctl.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>

#include "../common/log.h"
#include "../common/param.h"
#include "../common/timers.h"

namespace ub {
namespace net {

namespace {
constexpr int kRxBatch = 32;
constexpr size_t kRxSlot = 65536 + 256;
constexpr uint64_t kLingerNs = 2000000000ull;

inline int32_t seq_diff(uint32_t a, uint32_t b) { return (int32_t)(a - b); }

// shift a kSackWords*64 bit little-endian bitmap right by t (0 < t <= kSackBits)
inline void shr_bits(uint64_t* w, int t) {
  const int ws = t / 64, bs = t % 64;
  for (int i = 0; i < kSackWords; ++i) {
    const int s = i + ws;
    uint64_t lo = s < kSackWords ? w[s] : 0, hi = s + 1 < kSackWords ? w[s + 1] : 0;
    w[i] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
  }
}
inline int trailing_ones(const uint64_t* w) {
  int n = 0;
  for (int i = 0; i < kSackWords; ++i) {
    if (~w[i] == 0) {
      n += 64;
      continue;
    }
    return n + __builtin_ctzll(~w[i]);
  }
  return n;
}
}  // namespace

EngineConfig EngineConfig::from_env() {
  EngineConfig c;
  c.bind_ip = param_load_str("NET_BIND_IP", c.bind_ip.c_str());
  c.paths = (int)param_load("NET_PATHS", c.paths);
  c.payload = (int)param_load("NET_PAYLOAD", c.payload);
  c.max_inflight = (int)param_load("NET_MAX_INFLIGHT", c.max_inflight);
  c.eager_max = (size_t)param_load("NET_EAGER_MAX", (int64_t)c.eager_max);
  c.rto_min_us = (int)param_load("NET_RTO_MIN_US", c.rto_min_us);
  c.rto_abort = (int)param_load("NET_RTO_ABORT", c.rto_abort);
  c.link_gbps = (double)param_load("NET_LINK_GBPS", (int64_t)c.link_gbps);
  c.swift_target_us = (double)param_load("NET_SWIFT_TARGET_US", (int64_t)c.swift_target_us);
  c.busy_poll = param_load("NET_BUSY_POLL", 0) != 0;
  c.drop_prob = (double)param_load("NET_DROP_PPM", 0) * 1e-6;
  const std::string cc = param_load_str("NET_CC", "swift");
  if (cc == "none") c.cc = CC_NONE;
  else if (cc == "timely") c.cc = CC_TIMELY;
  else if (cc == "eqds") c.cc = CC_EQDS;
  else c.cc = CC_SWIFT;
  return c;
}

std::vector<std::pair<std::string, std::string>> list_interfaces() {
  std::vector<std::pair<std::string, std::string>> out, lo;
  ifaddrs* ifa = nullptr;
  if (getifaddrs(&ifa) != 0) return out;
  const std::string want = param_load_str("NET_IFNAME", "");
  for (ifaddrs* p = ifa; p; p = p->ifa_next) {
    if (!p->ifa_addr || p->ifa_addr->sa_family != AF_INET || !(p->ifa_flags & IFF_UP)) continue;
    char ip[INET_ADDRSTRLEN];
    inet_ntop(AF_INET, &reinterpret_cast<sockaddr_in*>(p->ifa_addr)->sin_addr, ip, sizeof(ip));
    const std::string name = p->ifa_name;
    if (!want.empty()) {
      // comma separated prefixes, like NCCL_SOCKET_IFNAME
      bool ok = false;
      size_t pos = 0;
      while (pos <= want.size()) {
        size_t e = want.find(',', pos);
        if (e == std::string::npos) e = want.size();
        const std::string tok = want.substr(pos, e - pos);
        if (!tok.empty() && name.compare(0, tok.size(), tok) == 0) ok = true;
        pos = e + 1;
      }
      if (!ok) continue;
      out.emplace_back(name, ip);
    } else if (p->ifa_flags & IFF_LOOPBACK) {
      lo.emplace_back(name, ip);
    } else if (name.compare(0, 6, "docker") != 0) {
      out.emplace_back(name, ip);
    }
  }
  freeifaddrs(ifa);
  if (out.empty()) out = lo;  // a box without a NIC still gets a (loopback) device
  return out;
}

namespace {
// MTU of the interface that owns `ip` (0 if unknown, e.g. 0.0.0.0)
int mtu_of_ip(const std::string& ip) {
  in_addr want{};
  if (inet_pton(AF_INET, ip.c_str(), &want) != 1 || want.s_addr == 0) return 0;
  ifaddrs* ifa = nullptr;
  if (getifaddrs(&ifa) != 0) return 0;
  int mtu = 0;
  for (ifaddrs* p = ifa; p; p = p->ifa_next) {
    if (!p->ifa_addr || p->ifa_addr->sa_family != AF_INET) continue;
    if (reinterpret_cast<sockaddr_in*>(p->ifa_addr)->sin_addr.s_addr != want.s_addr) continue;
    ifreq r{};
    snprintf(r.ifr_name, sizeof(r.ifr_name), "%s", p->ifa_name);
    int fd = socket(AF_INET, SOCK_DGRAM, 0);
    if (fd >= 0) {
      if (ioctl(fd, SIOCGIFMTU, &r) == 0) mtu = r.ifr_mtu;
      close(fd);
    }
    break;
  }
  freeifaddrs(ifa);
  return mtu;
}
}  // namespace

namespace {
// "0-3,8,10-11" -> cpu set
bool parse_cpulist(const std::string& txt, cpu_set_t* set) {
  CPU_ZERO(set);
  bool any = false;
  size_t pos = 0;
  while (pos < txt.size()) {
    size_t e = txt.find(',', pos);
    if (e == std::string::npos) e = txt.size();
    const std::string tok = txt.substr(pos, e - pos);
    int a = -1, b = -1;
    if (sscanf(tok.c_str(), "%d-%d", &a, &b) == 2 && a >= 0 && b >= a) {
      for (int c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET(c, set), any = true;
    } else if (sscanf(tok.c_str(), "%d", &a) == 1 && a >= 0 && a < CPU_SETSIZE) {
      CPU_SET(a, set), any = true;
    }
    pos = e + 1;
  }
  return any;
}

// Engine threads belong on the cores next to their NIC (the reference pins its engine threads per NIC as well):
// UCCL_B200_NET_CPUS="8-15" wins; otherwise the NIC's own local_cpulist from sysfs; otherwise no pinning.
void pin_engine_thread(std::thread& t, const std::string& bind_ip) {
  std::string list = param_load_str("NET_CPUS", "");
  if (list.empty()) {
    in_addr want{};
    if (inet_pton(AF_INET, bind_ip.c_str(), &want) == 1 && want.s_addr != 0) {
      ifaddrs* ifa = nullptr;
      if (getifaddrs(&ifa) == 0) {
        for (ifaddrs* p = ifa; p; p = p->ifa_next) {
          if (!p->ifa_addr || p->ifa_addr->sa_family != AF_INET) continue;
          if (reinterpret_cast<sockaddr_in*>(p->ifa_addr)->sin_addr.s_addr != want.s_addr) continue;
          char path[256];
          snprintf(path, sizeof(path), "/sys/class/net/%s/device/local_cpulist", p->ifa_name);
          if (FILE* f = fopen(path, "r")) {
            char buf[512] = {0};
            if (fgets(buf, sizeof(buf), f)) list = buf;
            fclose(f);
          }
          break;
        }
        freeifaddrs(ifa);
      }
    }
  }
  cpu_set_t set;
  if (list.empty() || !parse_cpulist(list, &set)) return;
  if (pthread_setaffinity_np(t.native_handle(), sizeof(set), &set) == 0)
    UB_INFO(SUB_NET, "net: engine thread of %s pinned to cpus %s", bind_ip.c_str(), list.c_str());
}
}  // namespace

// ------------------------------------------------------------------------------------------- setup
Engine::Engine(const EngineConfig& cfg) : cfg_(cfg), pacer_(cc::EqdsConfig()), rng_(std::random_device{}()) {
  cfg_.paths = std::max(1, std::min(cfg_.paths, kMaxPaths));
  cfg_.payload = std::max(256, std::min(cfg_.payload, 60000));
  if (param_load("NET_PAYLOAD_CLAMP", 1) != 0) {
    // keep every datagram inside the NIC's MTU: no IP fragmentation (one lost fragment loses the whole datagram)
    // and UDP GSO stays usable (its segments may not be fragmented)
    const int mtu = mtu_of_ip(cfg_.bind_ip);
    const int room = mtu - 28 - (int)sizeof(PktHdr);
    if (mtu > 0 && room >= 256 && cfg_.payload > room) {
      UB_INFO(SUB_NET, "net: payload %d -> %d to fit the %d-byte MTU of %s", cfg_.payload, room / 64 * 64, mtu, cfg_.bind_ip.c_str());
      cfg_.payload = room / 64 * 64;
    }
  }
  cfg_.max_inflight = std::max(4, std::min(cfg_.max_inflight, kSackBits - 8));
  drop_prob_.store(cfg_.drop_prob);
  cc::EqdsConfig ec;
  ec.link_gbps = cfg_.link_gbps;
  ec.credit_bytes = (uint32_t)cfg_.payload * 8;
  ec.max_backlog_credits = 8;
  pacer_ = cc::EqdsPacer(ec);
  for (int i = 0; i < kMaxPaths; ++i) socks_[i] = -1, ports_[i] = 0;
  gro_ok_ = param_load("NET_GRO", 1) != 0;
  epfd_ = epoll_create1(EPOLL_CLOEXEC);
  evfd_ = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
  UB_CHECK(epfd_ >= 0 && evfd_ >= 0, "net: epoll/eventfd: %s", strerror(errno));
  epoll_event ev{};
  ev.events = EPOLLIN;
  ev.data.u32 = 0xffffffffu;
  epoll_ctl(epfd_, EPOLL_CTL_ADD, evfd_, &ev);
  for (int i = 0; i < cfg_.paths; ++i) {
    int s = socket(AF_INET, SOCK_DGRAM | SOCK_NONBLOCK | SOCK_CLOEXEC, 0);
    UB_CHECK(s >= 0, "net: socket: %s", strerror(errno));
    int sz = cfg_.sockbuf_bytes;
    setsockopt(s, SOL_SOCKET, SO_RCVBUF, &sz, sizeof(sz));
    setsockopt(s, SOL_SOCKET, SO_SNDBUF, &sz, sizeof(sz));
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_port = 0;
    UB_CHECK(inet_pton(AF_INET, cfg_.bind_ip.c_str(), &a.sin_addr) == 1, "net: bad bind ip '%s'", cfg_.bind_ip.c_str());
    UB_CHECK(bind(s, reinterpret_cast<sockaddr*>(&a), sizeof(a)) == 0, "net: bind %s: %s", cfg_.bind_ip.c_str(),
             strerror(errno));
    socklen_t al = sizeof(a);
    getsockname(s, reinterpret_cast<sockaddr*>(&a), &al);
    if (gro_ok_) {
      int one = 1;
      if (setsockopt(s, IPPROTO_UDP, 104 /* UDP_GRO */, &one, sizeof(one)) != 0) gro_ok_ = false;
    }
    socks_[i] = s;
    ports_[i] = ntohs(a.sin_port);
    ev.data.u32 = (uint32_t)i;
    epoll_ctl(epfd_, EPOLL_CTL_ADD, s, &ev);
  }
  rx_buf_.resize((size_t)kRxBatch * kRxSlot);
  txb_.resize((size_t)cfg_.paths);
  gso_ok_ = param_load("NET_GSO", 1) != 0;
  UB_INFO(SUB_NET, "net engine up: %s paths=%d port0=%u payload=%d cc=%d", cfg_.bind_ip.c_str(), cfg_.paths, ports_[0],
          cfg_.payload, cfg_.cc);
  thr_ = std::thread([this] { run(); });
  pin_engine_thread(thr_, cfg_.bind_ip);
}

void Engine::shutdown(int linger_ms) {
  if (linger_ms < 0) linger_ms = (int)param_load("NET_LINGER_MS", 2000);
  {
    std::lock_guard<std::mutex> lk(mu_);
    if (shut_) return;
    shut_ = true;
    for (auto& kv : flows_) {
      const int st = kv.second->state.load();
      if (st == FL_ESTABLISHED || st == FL_SYN_SENT) {
        Cmd c;
        c.op = 4;
        c.flow = kv.first;
        cmds_.push_back(c);
      }
    }
  }
  wake();
  const uint64_t deadline = now_ns() + (uint64_t)linger_ms * 1000000ull;
  for (;;) {
    bool all_done = true;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto& kv : flows_) {
        const int st = kv.second->state.load();
        if (st == FL_ERROR) continue;
        if (st != FL_CLOSED) all_done = false;  // CLOSED = our FIN is acknowledged (or we gave up on the peer)
      }
    }
    const uint64_t now = now_ns();
    // quiet period: a peer whose last data (or FIN) we received may not have our ACK yet -- it retransmits within its
    // RTO, and every such packet restarts this timer.  We do not wait for the peer to close its side: whatever it
    // sends after our close would not be delivered to anybody anyway.
    if (all_done && now - last_rx_ns_.load() > std::max<uint64_t>(3ull * (uint64_t)cfg_.rto_min_us * 1000ull, 20000000ull)) break;
    if (now > deadline) break;
    std::this_thread::sleep_for(std::chrono::microseconds(200));
  }
}

Engine::~Engine() {
  shutdown(-1);
  stop_.store(true);
  wake();
  if (thr_.joinable()) thr_.join();
  {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& kv : flows_) {
      Flow& f = *kv.second;
      if (f.state.load() != FL_ERROR) fail_flow(f, nullptr);
    }
    for (auto& c : cmds_)
      if (c.req) complete(c.req, 0, 1);
    cmds_.clear();
  }
  for (int i = 0; i < cfg_.paths; ++i)
    if (socks_[i] >= 0) close(socks_[i]);
  if (epfd_ >= 0) close(epfd_);
  if (evfd_ >= 0) close(evfd_);
}

void Engine::wake() {
  uint64_t one = 1;
  ssize_t r = write(evfd_, &one, sizeof(one));
  (void)r;
}

std::shared_ptr<Engine::Flow> Engine::find(uint32_t id) const {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = flows_.find(id);
  return it == flows_.end() ? nullptr : it->second;
}

// ----------------------------------------------------------------------------------- app-side API
uint32_t Engine::listen() {
  std::lock_guard<std::mutex> lk(mu_);
  const uint32_t id = next_listen_++;
  listeners_[id];
  return id;
}

void Engine::close_listen(uint32_t listen_id) {
  std::lock_guard<std::mutex> lk(mu_);
  listeners_.erase(listen_id);
}

uint32_t Engine::connect_async(const std::string& ip, uint16_t port, uint32_t listen_id) {
  auto f = std::make_shared<Flow>();
  f->id = next_flow_++;
  f->listen_id = listen_id;
  UB_CHECK(inet_pton(AF_INET, ip.c_str(), &f->peer_ip) == 1, "net: bad peer ip '%s'", ip.c_str());
  f->syn_port = port;
  const uint64_t now = now_ns();
  f->syn_next_ns = now;
  f->syn_deadline_ns = now + (uint64_t)cfg_.connect_timeout_ms * 1000000ull;
  {
    std::lock_guard<std::mutex> lk(mu_);
    f->nonce = (((uint64_t)std::random_device{}() << 32) ^ now) | 1;  // rng_ belongs to the engine thread
    const int64_t pin = param_load("NET_ISN", -1);  // tests pin it just below the 32-bit wrap
    f->isn = pin >= 0 ? (uint32_t)pin : (uint32_t)std::random_device{}();
    f->snd_nxt = f->snd_una = f->isn;
    f->state.store(FL_SYN_SENT);
    flows_[f->id] = f;
    active_dirty_ = true;
  }
  wake();
  return f->id;
}

int Engine::flow_state(uint32_t flow) const {
  auto f = find(flow);
  return f ? f->state.load(std::memory_order_acquire) : 0;
}

bool Engine::accept_nb(uint32_t listen_id, uint32_t* flow) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = listeners_.find(listen_id);
  if (it == listeners_.end() || it->second.ready.empty()) return false;
  *flow = it->second.ready.front();
  it->second.ready.pop_front();
  return true;
}

uint32_t Engine::connect(const std::string& ip, uint16_t port, uint32_t listen_id, int timeout_ms) {
  const uint32_t id = connect_async(ip, port, listen_id);
  const uint64_t t0 = now_ns();
  for (;;) {
    const int st = flow_state(id);
    if (st == FL_ESTABLISHED) return id;
    UB_CHECK(st == FL_SYN_SENT, "net: connect to %s:%u failed (state %d)", ip.c_str(), port, st);
    UB_CHECK(timeout_ms < 0 || now_ns() - t0 < (uint64_t)timeout_ms * 1000000ull, "net: connect to %s:%u timed out",
             ip.c_str(), port);
    std::this_thread::sleep_for(std::chrono::microseconds(100));
  }
}

uint32_t Engine::accept(uint32_t listen_id, int timeout_ms) {
  const uint64_t t0 = now_ns();
  uint32_t id = 0;
  while (!accept_nb(listen_id, &id)) {
    UB_CHECK(timeout_ms < 0 || now_ns() - t0 < (uint64_t)timeout_ms * 1000000ull, "net: accept timed out");
    std::this_thread::sleep_for(std::chrono::microseconds(100));
  }
  return id;
}

void Engine::close_flow(uint32_t flow) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    Cmd c;
    c.op = 4;
    c.flow = flow;
    cmds_.push_back(c);
  }
  wake();
}

Request* Engine::send_async(uint32_t flow, const void* data, size_t bytes) {
  Request* r = new Request();
  r->flow = flow;
  r->is_send = true;
  {
    std::lock_guard<std::mutex> lk(mu_);
    Cmd c;
    c.op = 2;
    c.flow = flow;
    c.req = r;
    c.ptr = const_cast<void*>(data);
    c.len = bytes;
    cmds_.push_back(c);
  }
  wake();
  return r;
}

Request* Engine::recv_async(uint32_t flow, void* data, size_t capacity) {
  Request* r = new Request();
  r->flow = flow;
  {
    std::lock_guard<std::mutex> lk(mu_);
    Cmd c;
    c.op = 3;
    c.flow = flow;
    c.req = r;
    c.ptr = data;
    c.len = capacity;
    cmds_.push_back(c);
  }
  wake();
  return r;
}

bool Engine::test(Request* r, size_t* bytes, int* err) {
  if (!r->done.load(std::memory_order_acquire)) return false;
  if (bytes) *bytes = r->bytes;
  if (err) *err = r->err.load(std::memory_order_relaxed);
  delete r;
  return true;
}

Engine::WaitResult Engine::wait3(Request* r, size_t* bytes, int timeout_ms) {
  const uint64_t t0 = now_ns();
  int err = 0;
  uint32_t spins = 0;
  while (!test(r, bytes, &err)) {
    if (timeout_ms >= 0 && now_ns() - t0 > (uint64_t)timeout_ms * 1000000ull) return WAIT_TIMEOUT;  // request stays live
    if (++spins > 2000) std::this_thread::sleep_for(std::chrono::microseconds(20));
    else std::this_thread::yield();
  }
  return err == 0 ? WAIT_OK : WAIT_ERROR;
}

void Engine::abort_flow(uint32_t flow) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    Cmd c;
    c.op = 5;
    c.flow = flow;
    cmds_.push_back(c);
  }
  wake();
}

bool Engine::cancel(Request* r, int grace_ms) {
  abort_flow(r->flow);  // fail_flow() completes every request of the flow, `r` included
  size_t bytes = 0;
  return wait3(r, &bytes, grace_ms) != WAIT_TIMEOUT;
}

bool Engine::wait(Request* r, size_t* bytes, int timeout_ms) {
  const WaitResult w = wait3(r, bytes, timeout_ms);
  if (w == WAIT_TIMEOUT) {
    if (!cancel(r)) UB_WARN("net: engine thread did not release a timed-out request (leaked, buffer must stay valid)");
    return false;
  }
  return w == WAIT_OK;
}

EngineStats Engine::stats() const {
  std::lock_guard<std::mutex> lk(st_mu_);
  return est_;
}

bool Engine::flow_stats(uint32_t flow, FlowStats* out) const {
  auto f = find(flow);
  if (!f) return false;
  std::lock_guard<std::mutex> lk(st_mu_);
  *out = f->st;
  out->state = f->state.load();
  return true;
}

void Engine::complete(Request* r, size_t bytes, int err) {
  if (!r) return;
  r->bytes = bytes;
  r->err.store(err, std::memory_order_relaxed);
  r->done.store(1, std::memory_order_release);
}

// ------------------------------------------------------------------------------------ engine loop
void Engine::run() {
  uint64_t last_stats = 0, last_busy_ns = 0;
  const uint64_t spin_ns = (uint64_t)param_load("NET_SPIN_US", 0) * 1000ull;
  const int pace_nap_us = (int)param_load("NET_PACE_NAP_US", 20);
  while (!stop_.load(std::memory_order_relaxed)) {
    bool busy = false;
    // commands + working set
    {
      std::vector<Cmd> cmds;
      {
        std::lock_guard<std::mutex> lk(mu_);
        cmds.swap(cmds_);
        if (active_dirty_) {
          active_.clear();
          index_.clear();
          for (auto& kv : flows_) {
            active_.push_back(kv.second);
            index_[kv.first] = kv.second.get();
          }
          active_dirty_ = false;
        }
      }
      for (auto& c : cmds) {
        busy = true;
        Flow* f = nullptr;
        {
          auto it = index_.find(c.flow);  // rebuilt above whenever the flow table changed
          if (it != index_.end()) f = it->second;
        }
        if (!f) {
          if (c.req) complete(c.req, 0, 1);
          continue;
        }
        if (c.op == 2) post_send(*f, c.req, c.ptr, c.len);
        else if (c.op == 3) post_recv(*f, c.req, c.ptr, c.len);
        else if (c.op == 4) {
          const int st = f->state.load();
          if (st == FL_ESTABLISHED) {
            f->fin_pending = true;
            f->state.store(FL_CLOSING);
            f->last_progress_ns = now_ns();
            f->close_start_ns = f->last_progress_ns;
          } else if (st == FL_SYN_SENT) {
            fail_flow(*f, nullptr);
          }
        } else if (c.op == 5) {
          if (f->state.load() != FL_ERROR) fail_flow(*f, "aborted by the application (request timed out)");
        }
      }
    }
    busy |= rx_poll();
    const uint64_t now = now_ns();
    timers(now);
    if (cfg_.cc == CC_EQDS) eqds_tick(now);
    bool pending = false;
    for (auto& sp : active_) {
      Flow& f = *sp;
      const int st = f.state.load(std::memory_order_relaxed);
      if (st == FL_ESTABLISHED || st == FL_CLOSING) {
        busy |= tx_pump(f, now);
        if (f.need_ack || f.credit_dirty) send_ack(f);
        if (f.snd_una != f.snd_nxt || f.tx_cursor < f.txq.size()) pending = true;
      } else if (st == FL_SYN_SENT) {
        pending = true;
      } else if (st == FL_CLOSED && f.need_ack) {
        send_ack(f);  // linger: late retransmissions of the peer still get their ACK
      }
    }
    flush_all();
    if (!held_.empty()) {
      release_held(now_ns());
      busy = true;  // keep the loop turning until the held datagrams have left
    }
    ++est_.loops;
    if (now - last_stats > 1000000ull) {
      last_stats = now;
      std::lock_guard<std::mutex> lk(st_mu_);
      // counters are single-writer (this thread) aligned words; readers tolerate slightly stale values
      for (auto& sp : active_) {
        Flow& f = *sp;
        f.st.srtt_us = f.srtt_us;
        f.st.min_rtt_us = f.min_rtt_us;
        f.st.cwnd = f.swift.cwnd();
        f.st.rate_gbps = f.timely.rate_gbps();
      }
      est_.flows = (int)active_.size();
    }
    if (busy) last_busy_ns = now;
    if (!busy && !cfg_.busy_poll) {
      // hybrid polling: keep spinning for a short while after the last activity -- request/response
      // patterns (ring steps, rendezvous RTR -> data) then never pay an epoll wake-up per hop
      if (now - last_busy_ns < spin_ns) continue;
      const bool eq = cfg_.cc == CC_EQDS && pacer_.active_senders() > 0;
      const int to = pending ? 1 : 50;
      // the credit pacer and the rate pacer need a fine-grained clock while there is work: nap instead of
      // sleeping in epoll.  (A hard spin is faster on an idle core but gets the thread throttled in 4 ms
      // quanta under a CPU quota; UCCL_B200_NET_BUSY_POLL=1 selects the spin.)
      if (eq || (cfg_.cc == CC_TIMELY && pending)) {
        std::this_thread::sleep_for(std::chrono::microseconds(pace_nap_us));
        continue;
      }
      ++est_.sleeps;
      epoll_event evs[8];
      const int n = epoll_wait(epfd_, evs, 8, to);
      for (int i = 0; i < n; ++i)
        if (evs[i].data.u32 == 0xffffffffu) {
          uint64_t v;
          ssize_t r = read(evfd_, &v, sizeof(v));
          (void)r;
        }
    }
  }
}

bool Engine::rx_poll() {
  bool any = false;
  mmsghdr msgs[kRxBatch];
  iovec iov[kRxBatch];
  sockaddr_in from[kRxBatch];
  alignas(cmsghdr) char ctrl[kRxBatch][CMSG_SPACE(sizeof(int))];
  for (int s = 0; s < cfg_.paths; ++s) {
    for (int round = 0; round < 4; ++round) {
      for (int i = 0; i < kRxBatch; ++i) {
        iov[i].iov_base = rx_buf_.data() + (size_t)i * kRxSlot;
        iov[i].iov_len = kRxSlot;
        memset(&msgs[i], 0, sizeof(msgs[i]));
        msgs[i].msg_hdr.msg_iov = &iov[i];
        msgs[i].msg_hdr.msg_iovlen = 1;
        msgs[i].msg_hdr.msg_name = &from[i];
        msgs[i].msg_hdr.msg_namelen = sizeof(sockaddr_in);
        if (gro_ok_) {
          msgs[i].msg_hdr.msg_control = ctrl[i];
          msgs[i].msg_hdr.msg_controllen = sizeof(ctrl[i]);
        }
      }
      const int n = recvmmsg(socks_[s], msgs, kRxBatch, MSG_DONTWAIT, nullptr);
      if (n <= 0) break;
      any = true;
      for (int i = 0; i < n; ++i) {
        uint8_t* buf = static_cast<uint8_t*>(iov[i].iov_base);
        size_t len = msgs[i].msg_len;
        // UDP GRO: the kernel hands over a train of equal-sized datagrams of one sender as one buffer and
        // reports the segment size in a control message; every segment starts with its own PktHdr
        size_t seg = 0;
        if (gro_ok_)
          for (cmsghdr* cm = CMSG_FIRSTHDR(&msgs[i].msg_hdr); cm; cm = CMSG_NXTHDR(&msgs[i].msg_hdr, cm))
            if (cm->cmsg_level == IPPROTO_UDP && cm->cmsg_type == 104 /* UDP_GRO */) {
              int v = 0;
              memcpy(&v, CMSG_DATA(cm), sizeof(v));
              seg = v > 0 ? (size_t)v : 0;
            }
        if (seg == 0 || seg >= len) {
          on_packet(s, from[i], buf, len);
        } else {
          for (size_t off = 0; off < len; off += seg) on_packet(s, from[i], buf + off, std::min(seg, len - off));
        }
      }
      if (n < kRxBatch) break;
    }
  }
  return any;
}

void Engine::on_packet(int sock_idx, const sockaddr_in& from, uint8_t* buf, size_t n) {
  last_rx_ns_.store(now_ns(), std::memory_order_relaxed);
  ++est_.rx_pkts;
  est_.rx_bytes += n;
  if (n < sizeof(PktHdr)) {
    ++est_.bad_pkts;
    return;
  }
  PktHdr h;
  memcpy(&h, buf, sizeof(h));
  if (h.magic != kMagic || sizeof(PktHdr) + h.len > n) {
    ++est_.bad_pkts;
    return;
  }
  const uint8_t* body = buf + sizeof(PktHdr);
  if (h.type == PKT_SYN || h.type == PKT_SYNACK) {
    if (h.len < sizeof(SynBody)) {
      ++est_.bad_pkts;
      return;
    }
    SynBody b;
    memcpy(&b, body, sizeof(b));
    if (h.type == PKT_SYN) on_syn(sock_idx, from, h, b);
    else on_synack(from, h, b);
    return;
  }
  Flow* f = nullptr;
  {
    auto it = index_.find(h.dst_flow);  // engine-thread index of active_; the lock-protected map is the slow path
    if (it != index_.end()) f = it->second;
  }
  std::shared_ptr<Flow> hold;
  if (!f) {
    hold = find(h.dst_flow);
    f = hold.get();
  }
  if (!f || f->peer_ip.s_addr != from.sin_addr.s_addr) {
    if (h.type == PKT_DATA) send_rst(sock_idx, from, 0);
    return;
  }
  const int st = f->state.load(std::memory_order_relaxed);
  if (h.type == PKT_RST) {
    if (st == FL_ESTABLISHED || st == FL_SYN_SENT || st == FL_CLOSING) fail_flow(*f, "reset by peer");
    return;
  }
  if (st == FL_SYN_SENT) return;  // SYNACK not seen yet: the peer will retransmit
  if (st == FL_ERROR) {
    if (h.type == PKT_DATA) send_rst(sock_idx, from, f->peer_flow);
    return;
  }
  if (h.type == PKT_DATA) {
    on_data(*f, h, body);
  } else if (h.type == PKT_ACK) {
    if (h.len < sizeof(AckBody)) return;
    AckBody b;
    memcpy(&b, body, sizeof(b));
    on_ack(*f, h, b);
  }
}

// ------------------------------------------------------------------------------------- handshake
void Engine::fill_syn_body(SynBody* b, const Flow& f) const {
  memset(b, 0, sizeof(*b));
  b->nonce = f.nonce;
  b->src_flow = f.id;
  b->listen_id = f.listen_id;
  b->npaths = (uint16_t)cfg_.paths;
  for (int i = 0; i < cfg_.paths; ++i) b->ports[i] = htons(ports_[i]);
}

void Engine::send_syn(Flow& f, bool synack, int sock_idx, const sockaddr_in* to) {
  PktHdr h{};
  h.magic = kMagic;
  h.type = synack ? PKT_SYNACK : PKT_SYN;
  h.dst_flow = synack ? f.peer_flow : 0;
  h.len = sizeof(SynBody);
  h.ts_ns = now_ns();
  h.seq = f.isn;  // our direction starts here; random so that stale datagrams of an earlier flow never match
  SynBody b;
  fill_syn_body(&b, f);
  sockaddr_in a{};
  if (to) {
    a = *to;
  } else {
    a.sin_family = AF_INET;
    a.sin_addr = f.peer_ip;
    a.sin_port = htons(f.syn_port);
  }
  raw_send(sock_idx, a, &h, sizeof(h), &b, sizeof(b));
}

void Engine::on_syn(int sock_idx, const sockaddr_in& from, const PktHdr& h, const SynBody& b) {
  std::shared_ptr<Flow> f;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto lit = listeners_.find(b.listen_id);
    if (lit == listeners_.end()) {
      f = nullptr;
    } else {
      const std::pair<uint64_t, uint64_t> key(((uint64_t)from.sin_addr.s_addr << 32) | b.src_flow, b.nonce);
      auto it = syn_index_.find(key);
      if (it != syn_index_.end()) {
        auto fit = flows_.find(it->second);
        if (fit != flows_.end()) f = fit->second;
      } else {
        f = std::make_shared<Flow>();
        f->id = next_flow_++;
        f->peer_flow = b.src_flow;
        f->nonce = b.nonce;
        f->listen_id = b.listen_id;
        f->peer_ip = from.sin_addr;
        f->npaths = std::max(1, std::min<int>(cfg_.paths, b.npaths));
        for (int i = 0; i < f->npaths; ++i) {
          f->peer_addr[i] = sockaddr_in{};
          f->peer_addr[i].sin_family = AF_INET;
          f->peer_addr[i].sin_addr = from.sin_addr;
          f->peer_addr[i].sin_port = b.ports[i];
        }
        f->rcv_nxt = h.seq;  // the client's initial sequence number
        {
          const int64_t pin = param_load("NET_ISN", -1);
          f->isn = pin >= 0 ? (uint32_t)pin : (uint32_t)rng_();
          f->snd_nxt = f->snd_una = f->isn;
        }
        f->rto_ns = (uint64_t)cfg_.rto_min_us * 1000ull;
        cc::SwiftConfig sc;
        sc.base_target_us = cfg_.swift_target_us;
        sc.max_cwnd = cfg_.max_inflight;
        f->swift = cc::Swift(sc);
        cc::TimelyConfig tc;
        tc.min_rtt_us = 50, tc.t_low_us = 200, tc.t_high_us = 2000, tc.link_gbps = cfg_.link_gbps;
        tc.add_step_gbps = cfg_.link_gbps / 40, tc.min_rate_gbps = cfg_.link_gbps / 1000;
        f->timely = cc::Timely(tc);
        f->state.store(FL_ESTABLISHED, std::memory_order_release);
        flows_[f->id] = f;
        syn_index_[key] = f->id;
        lit->second.ready.push_back(f->id);
        active_.push_back(f);
        index_[f->id] = f.get();
      }
    }
  }
  if (!f) {
    send_rst(sock_idx, from, b.src_flow);
    return;
  }
  send_syn(*f, true, 0, &from);
}

void Engine::on_synack(const sockaddr_in& from, const PktHdr& h, const SynBody& b) {
  std::shared_ptr<Flow> f = find(h.dst_flow);
  if (!f || f->state.load() != FL_SYN_SENT || b.nonce != f->nonce || from.sin_addr.s_addr != f->peer_ip.s_addr) return;
  f->peer_flow = b.src_flow;
  f->rcv_nxt = h.seq;  // the server's initial sequence number
  f->npaths = std::max(1, std::min<int>(cfg_.paths, b.npaths));
  for (int i = 0; i < f->npaths; ++i) {
    f->peer_addr[i] = sockaddr_in{};
    f->peer_addr[i].sin_family = AF_INET;
    f->peer_addr[i].sin_addr = f->peer_ip;
    f->peer_addr[i].sin_port = b.ports[i];
  }
  f->rto_ns = (uint64_t)cfg_.rto_min_us * 1000ull;
  cc::SwiftConfig sc;
  sc.base_target_us = cfg_.swift_target_us;
  sc.max_cwnd = cfg_.max_inflight;
  f->swift = cc::Swift(sc);
  cc::TimelyConfig tc;
  tc.min_rtt_us = 50, tc.t_low_us = 200, tc.t_high_us = 2000, tc.link_gbps = cfg_.link_gbps;
  tc.add_step_gbps = cfg_.link_gbps / 40, tc.min_rate_gbps = cfg_.link_gbps / 1000;
  f->timely = cc::Timely(tc);
  f->state.store(FL_ESTABLISHED, std::memory_order_release);
}

void Engine::send_rst(int sock_idx, const sockaddr_in& to, uint32_t dst_flow) {
  if (dst_flow == 0) return;  // we do not know the peer's flow id: stay silent, its RTO will give up
  PktHdr h{};
  h.magic = kMagic;
  h.type = PKT_RST;
  h.dst_flow = dst_flow;
  raw_send(sock_idx, to, &h, sizeof(h), nullptr, 0);
}

// ------------------------------------------------------------------------------------- raw output
void Engine::raw_send(int path, const sockaddr_in& to, const void* hdr, size_t hlen, const void* body, size_t blen) {
  const double dp = drop_prob_.load(std::memory_order_relaxed);
  if (dp > 0.0 && std::uniform_real_distribution<double>(0.0, 1.0)(rng_) < dp) {
    ++est_.dropped_tx;
    return;
  }
  (void)hlen;
  if (path == path_drop_idx_.load(std::memory_order_relaxed)) {
    const double pp = path_drop_prob_.load(std::memory_order_relaxed);
    if (pp > 0.0 && std::uniform_real_distribution<double>(0.0, 1.0)(rng_) < pp) {
      ++est_.dropped_tx;
      return;
    }
  }
  const double rp = reorder_prob_.load(std::memory_order_relaxed);
  if (rp > 0.0 && std::uniform_real_distribution<double>(0.0, 1.0)(rng_) < rp) {
    Held h;
    h.release_ns = now_ns() + (uint64_t)reorder_delay_us_.load(std::memory_order_relaxed) * 1000ull;
    h.path = path;
    h.to = to;
    h.bytes.resize(sizeof(PktHdr) + blen);
    memcpy(h.bytes.data(), hdr, sizeof(PktHdr));
    if (blen) memcpy(h.bytes.data() + sizeof(PktHdr), body, blen);
    held_.push_back(std::move(h));
    return;
  }
  TxBatch& b = txb_[path];
  if (b.n == kTxBatch) flush_path(path);
  TxSlot& s = b.slot[b.n++];
  memcpy(&s.hdr, hdr, sizeof(PktHdr));
  s.to = to;
  s.blen = (uint32_t)blen;
  if (s.hdr.type == PKT_DATA) {
    s.payload = body;  // user buffer: stays valid until the packet is acknowledged
  } else {
    if (blen) memcpy(s.body, body, std::min(blen, sizeof(s.body)));
    s.payload = s.body;
  }
}

void Engine::flush_path(int path) {
  TxBatch& b = txb_[path];
  if (b.n == 0) return;
  // Build the messages of one sendmmsg.  With UDP GSO (UDP_SEGMENT) a run of equal-sized DATA packets to
  // the same destination leaves as ONE super-datagram whose iovec is hdr0,payload0,hdr1,payload1,...: the
  // kernel cuts it every gso_size bytes, so the stack is traversed once per run instead of once per packet.
  mmsghdr msgs[kTxBatch];
  iovec iov[kTxBatch * 2];
  alignas(cmsghdr) char ctrl[kTxBatch][CMSG_SPACE(sizeof(uint16_t))];
  int run_len[kTxBatch];
  int nm = 0, i = 0;
  while (i < b.n) {
    TxSlot& s0 = b.slot[i];
    int jn = i + 1;
    const size_t seg = sizeof(PktHdr) + s0.blen;
    if (gso_ok_ && s0.hdr.type == PKT_DATA && s0.blen > 0) {
      size_t total = seg;
      while (jn < b.n && jn - i < 64) {
        TxSlot& sj = b.slot[jn];
        if (sj.hdr.type != PKT_DATA || sj.to.sin_port != s0.to.sin_port || sj.to.sin_addr.s_addr != s0.to.sin_addr.s_addr) break;
        if (sj.blen > s0.blen || sj.blen == 0 || total + sizeof(PktHdr) + sj.blen > 65000) break;
        total += sizeof(PktHdr) + sj.blen;
        ++jn;
        if (sj.blen < s0.blen) break;  // a shorter segment may only be the last one
      }
    }
    iovec* v = &iov[2 * i];
    for (int k = i; k < jn; ++k) {
      TxSlot& s = b.slot[k];
      v[2 * (k - i)].iov_base = &s.hdr;
      v[2 * (k - i)].iov_len = sizeof(PktHdr);
      v[2 * (k - i) + 1].iov_base = const_cast<void*>(s.payload);
      v[2 * (k - i) + 1].iov_len = s.blen;
    }
    memset(&msgs[nm], 0, sizeof(msgs[nm]));
    msgs[nm].msg_hdr.msg_name = &s0.to;
    msgs[nm].msg_hdr.msg_namelen = sizeof(s0.to);
    msgs[nm].msg_hdr.msg_iov = v;
    msgs[nm].msg_hdr.msg_iovlen = (jn - i == 1 && s0.blen == 0) ? 1 : (size_t)(2 * (jn - i));
    if (jn - i > 1) {
      msgs[nm].msg_hdr.msg_control = ctrl[nm];
      msgs[nm].msg_hdr.msg_controllen = sizeof(ctrl[nm]);
      cmsghdr* cm = CMSG_FIRSTHDR(&msgs[nm].msg_hdr);
      cm->cmsg_level = IPPROTO_UDP;  // == SOL_UDP
      cm->cmsg_type = 103;  // UDP_SEGMENT
      cm->cmsg_len = CMSG_LEN(sizeof(uint16_t));
      const uint16_t gs = (uint16_t)seg;
      memcpy(CMSG_DATA(cm), &gs, sizeof(gs));
    }
    run_len[nm] = jn - i;
    ++nm;
    i = jn;
  }
  int sent_msgs = 0;
  while (sent_msgs < nm) {
    const int r = sendmmsg(socks_[path], msgs + sent_msgs, (unsigned)(nm - sent_msgs), MSG_DONTWAIT);
    if (r <= 0) {
      const bool gso_refused = errno == EINVAL || errno == EIO || errno == EOPNOTSUPP || errno == ENOPROTOOPT ||
                               errno == EMSGSIZE;
      if (gso_ok_ && r < 0 && run_len[sent_msgs] > 1 && gso_refused) {
        // no UDP GSO on this kernel / device, or segments larger than the path MTU (EMSGSIZE: GSO segments may not
        // be IP-fragmented): fall back to one datagram per packet for good
        gso_ok_ = false;
        UB_INFO(SUB_NET, "net: UDP_SEGMENT unavailable (%s): GSO off", strerror(errno));
        int done = 0;
        for (int m = 0; m < sent_msgs; ++m) done += run_len[m];
        memmove(&b.slot[0], &b.slot[done], sizeof(TxSlot) * (size_t)(b.n - done));
        b.n -= done;
        for (int k = 0; k < b.n; ++k)  // control packets reference their own slot's body: re-point after the move
          if (b.slot[k].hdr.type != PKT_DATA) b.slot[k].payload = b.slot[k].body;
        est_.tx_pkts += (uint64_t)done;
        flush_path(path);
        return;
      }
      break;  // full socket buffer: the rest counts as dropped; the reliability layer repairs it
    }
    sent_msgs += r;
  }
  int sent = 0;
  for (int m = 0; m < sent_msgs; ++m) sent += run_len[m];
  for (int k = 0; k < sent; ++k) est_.tx_bytes += sizeof(PktHdr) + b.slot[k].blen;
  est_.tx_pkts += (uint64_t)sent;
  est_.dropped_tx += (uint64_t)(b.n - sent);
  b.n = 0;
}

void Engine::release_held(uint64_t now) {
  while (!held_.empty() && held_.front().release_ns <= now) {
    Held& h = held_.front();
    (void)::sendto(socks_[h.path], h.bytes.data(), h.bytes.size(), MSG_DONTWAIT, reinterpret_cast<sockaddr*>(&h.to), sizeof(h.to));
    ++est_.tx_pkts;
    held_.pop_front();
  }
}

void Engine::flush_all() {
  for (int p = 0; p < cfg_.paths; ++p) flush_path(p);
}

// ------------------------------------------------------------------------------- message posting
void Engine::post_send(Flow& f, Request* r, const void* ptr, size_t len) {
  const int st = f.state.load();
  if (st != FL_ESTABLISHED && st != FL_SYN_SENT) {
    complete(r, 0, 1);
    return;
  }
  TxMsg* m = new TxMsg();
  m->req = r;
  m->ptr = static_cast<const uint8_t*>(ptr);
  m->len = len;
  m->id = f.next_tx_msg++;
  f.txq.push_back(m);
}

void Engine::post_recv(Flow& f, Request* r, void* ptr, size_t cap) {
  const int st = f.state.load();
  if (st != FL_ESTABLISHED && st != FL_SYN_SENT) {
    complete(r, 0, 1);
    return;
  }
  const uint32_t id = f.rx_posted++;
  RxMsg m;
  m.req = r;
  m.ptr = static_cast<uint8_t*>(ptr);
  m.cap = cap;
  auto it = f.unexpected.find(id);
  if (it != f.unexpected.end()) {
    Unexpected& u = it->second;
    m.total = u.total;
    m.have_total = true;
    m.got = u.got;
    if (u.total > cap) m.overflow = true;
    else if (u.total) memcpy(m.ptr, u.buf.data(), u.total);
    f.unexpected.erase(it);
  }
  if (m.have_total && m.got >= m.total) {
    complete(r, m.total, m.overflow ? 2 : 0);
    m.req = nullptr;
  }
  f.rxq.push_back(m);
  while (!f.rxq.empty() && f.rxq.front().req == nullptr) {
    f.rxq.pop_front();
    ++f.rx_base;
  }
  if (f.peer_fin.load(std::memory_order_relaxed) && !f.rxq.empty()) {
    for (auto& q : f.rxq)
      if (q.req) complete(q.req, 0, 3), q.req = nullptr;
    while (!f.rxq.empty()) f.rxq.pop_front(), ++f.rx_base;
  }
  f.rtr_pending = true;
}

// ------------------------------------------------------------------------------------ receive side
void Engine::on_data(Flow& f, const PktHdr& h, const uint8_t* payload) {
  ++f.st.rx_pkts;
  f.need_ack = true;
  f.echo_ts = h.ts_ns;
  f.echo_path = h.path;
  if (cfg_.cc == CC_EQDS && h.aux > f.demand_seen) {
    pacer_.add_demand(f.id, h.aux - f.demand_seen);
    f.demand_seen = h.aux;
  }
  const int32_t d = seq_diff(h.seq, f.rcv_nxt);
  if (d < 0) {
    ++f.st.rx_dup;
    return;
  }
  if (d >= kSackBits) return;  // beyond the window: the sender never does this unless state is stale
  uint64_t& w = f.rx_bits[d / 64];
  const uint64_t bit = 1ull << (d % 64);
  if (w & bit) {
    ++f.st.rx_dup;
    return;
  }
  w |= bit;
  deliver_frame(f, h, payload);
  const int t = trailing_ones(f.rx_bits);
  if (t > 0) {
    shr_bits(f.rx_bits, t);
    f.rcv_nxt += (uint32_t)t;
  }
  if (f.have_fin && !f.peer_fin.load(std::memory_order_relaxed) && seq_diff(f.rcv_nxt, f.fin_seq) > 0) apply_peer_fin(f);
}

void Engine::apply_peer_fin(Flow& f) {
  f.peer_fin.store(true);
  for (auto& q : f.rxq)
    if (q.req) complete(q.req, 0, 3), q.req = nullptr;
  while (!f.rxq.empty()) f.rxq.pop_front(), ++f.rx_base;
}

void Engine::deliver_frame(Flow& f, const PktHdr& h, const uint8_t* payload) {
  if (h.kind == FR_RTR) {
    if (seq_diff(h.msg_id, f.peer_posted) > 0) f.peer_posted = h.msg_id;
    return;
  }
  if (h.kind == FR_FIN) {
    // frames are delivered out of order: the FIN only takes effect once everything before it has arrived
    f.have_fin = true;
    f.fin_seq = h.seq;
    return;
  }
  if (h.kind != FR_MSG) return;
  f.st.rx_bytes += h.len;
  if (cfg_.cc == CC_EQDS) pacer_.on_data(f.id, h.len);
  const int32_t idx = seq_diff(h.msg_id, f.rx_base);
  if (idx < 0) return;
  if ((size_t)idx < f.rxq.size()) {
    RxMsg& m = f.rxq[(size_t)idx];
    if (!m.req) return;
    if (!m.have_total) {
      m.total = h.msg_len;
      m.have_total = true;
      if (m.total > m.cap) m.overflow = true;
    }
    if (!m.overflow && h.len && h.offset <= m.total && h.len <= m.total - h.offset) memcpy(m.ptr + h.offset, payload, h.len);
    m.got += h.len;
    if (m.got >= m.total) {
      complete(m.req, m.total, m.overflow ? 2 : 0);
      m.req = nullptr;
      while (!f.rxq.empty() && f.rxq.front().req == nullptr) {
        f.rxq.pop_front();
        ++f.rx_base;
      }
    }
    return;
  }
  // receive not posted yet: only eager messages get here
  if (h.msg_len > (64ull << 20) || f.unexpected.size() > (size_t)(8 * cfg_.eager_ahead + 64)) {
    fail_flow(f, "unexpected-message buffer limits exceeded (protocol violation)");
    return;
  }
  auto it = f.unexpected.find(h.msg_id);
  if (it == f.unexpected.end()) {
    Unexpected u;
    u.total = h.msg_len;
    u.buf.resize(h.msg_len);
    it = f.unexpected.emplace(h.msg_id, std::move(u)).first;
    ++f.st.unexpected_msgs;
  }
  Unexpected& u = it->second;
  if (h.len && h.offset <= u.total && h.len <= u.total - h.offset) memcpy(u.buf.data() + h.offset, payload, h.len);
  u.got += h.len;
}

void Engine::send_ack(Flow& f) {
  PktHdr h{};
  h.magic = kMagic;
  h.type = PKT_ACK;
  const int path = f.npaths > 0 ? f.echo_path % f.npaths : 0;
  h.path = (uint16_t)path;
  h.dst_flow = f.peer_flow;
  h.seq = f.rcv_nxt;
  h.ts_ns = f.echo_ts;
  h.msg_id = f.rx_posted;
  h.len = sizeof(AckBody);
  h.aux = f.grant_cum;
  AckBody b{};
  for (int i = 0; i < kSackWords; ++i) b.sack[i] = f.rx_bits[i];
  b.echo_path = f.echo_path;
  b.dup_cum = (uint32_t)f.st.rx_dup;
  raw_send(path, f.peer_addr[path], &h, sizeof(h), &b, sizeof(b));
  ++f.st.acks_tx;
  f.need_ack = false;
  f.credit_dirty = false;
  f.echo_ts = 0;
}

void Engine::eqds_tick(uint64_t now) {
  auto grants = pacer_.tick((double)now * 1e-3);
  for (auto& g : grants)
    for (auto& sp : active_)
      if (sp->id == g.first) {
        sp->grant_cum += g.second;
        sp->credit_dirty = true;
        break;
      }
}

// --------------------------------------------------------------------------------------- send side
void Engine::mark_acked(Flow& f, TxPkt& p, uint64_t now) {
  (void)now;
  p.acked = true;
  if (!p.lost) {
    if (f.inflight) --f.inflight;
    if (f.path[p.path].inflight) --f.path[p.path].inflight;
  }
  p.lost = false;
  f.path[p.path].loss_streak = 0;
  if (p.ts_send > f.newest_acked_send_ts) f.newest_acked_send_ts = p.ts_send;
  TxMsg* m = p.msg;
  p.msg = nullptr;
  if (m) {
    m->acked += p.len;
    ++m->pkts_acked;
    if (m->all_queued && m->pkts_acked == m->pkts_out) {
      complete(m->req, m->len, 0);
      m->req = nullptr;
      while (!f.txq.empty() && f.txq.front()->req == nullptr && f.txq.front()->all_queued &&
             f.txq.front()->pkts_acked == f.txq.front()->pkts_out) {
        delete f.txq.front();
        f.txq.pop_front();
        if (f.tx_cursor) --f.tx_cursor;
      }
    }
  }
}

void Engine::on_ack(Flow& f, const PktHdr& h, const AckBody& b) {
  ++f.st.acks_rx;
  const uint64_t now = now_ns();
  if (seq_diff(h.msg_id, f.peer_posted) > 0) f.peer_posted = h.msg_id;
  if (h.aux > f.credit_cum) f.credit_cum = h.aux;
  if ((int32_t)(b.dup_cum - f.peer_dup_seen) > 0) {
    // the receiver got packets twice: some of our retransmissions were not needed, i.e. the network reorders by
    // more than the current window -- widen it (RACK's DSACK adaptation); it decays again after a quiet while
    f.peer_dup_seen = b.dup_cum;
    if (f.reo_mult < 16) ++f.reo_mult;
    f.reo_decay_ns = now + 16ull * (uint64_t)(std::max(f.srtt_us, 100.0) * 1e3);
  } else if (f.reo_mult > 1 && now > f.reo_decay_ns) {
    --f.reo_mult;
    f.reo_decay_ns = now + 16ull * (uint64_t)(std::max(f.srtt_us, 100.0) * 1e3);
  }
  const uint32_t cum = h.seq;
  const int32_t adv = seq_diff(cum, f.snd_una);
  if (adv < 0 || seq_diff(cum, f.snd_nxt) > 0) return;  // stale or nonsensical
  int newly = 0;
  for (int32_t i = 0; i < adv; ++i) {
    TxPkt& p = f.ring[(f.snd_una + (uint32_t)i) % kTxRing];
    if (p.in_use) {
      if (!p.acked) mark_acked(f, p, now), ++newly;
      p.in_use = false;
    }
  }
  f.snd_una = cum;
  for (int i = 1; i < kSackBits; ++i) {
    if (!(b.sack[i / 64] >> (i % 64) & 1)) continue;
    const uint32_t s = cum + (uint32_t)i;
    if (seq_diff(s, f.snd_nxt) >= 0) break;
    TxPkt& p = f.ring[s % kTxRing];
    if (p.in_use && p.seq == s && !p.acked) mark_acked(f, p, now), ++newly;
  }
  if (h.ts_ns != 0 && now > h.ts_ns) {
    const double rtt = (double)(now - h.ts_ns) * 1e-3;
    if (f.srtt_us == 0) {
      f.srtt_us = rtt;
      f.rttvar_us = rtt / 2;
      f.min_rtt_us = rtt;
    } else {
      f.rttvar_us = 0.75 * f.rttvar_us + 0.25 * std::abs(rtt - f.srtt_us);
      f.srtt_us = 0.875 * f.srtt_us + 0.125 * rtt;
      f.min_rtt_us = std::min(f.min_rtt_us, rtt);
    }
    PathState& ps = f.path[b.echo_path % kMaxPaths];
    ps.srtt_us = ps.srtt_us == 0 ? rtt : 0.875 * ps.srtt_us + 0.125 * rtt;
    if (newly) {
      if (cfg_.cc == CC_SWIFT) f.swift.on_ack(rtt, newly, (double)now * 1e-3, f.srtt_us);
      else if (cfg_.cc == CC_TIMELY) f.timely.on_rtt(rtt);
    }
  }
  if (newly) {
    f.tlp_fired = false;
    f.rto_count = 0;
    const double rto_us = std::max<double>(cfg_.rto_min_us, f.srtt_us + 4 * f.rttvar_us);
    f.rto_ns = (uint64_t)(std::min<double>(rto_us, cfg_.rto_max_us) * 1e3);
    f.last_progress_ns = now;
  }
  detect_loss(f, now);
}

void Engine::note_path_loss(Flow& f, int path, uint64_t now) {
  PathState& ps = f.path[path];
  if (++ps.loss_streak < 8 || now < ps.banned_until_ns) return;
  // eight losses in a row with no ACK in between: treat the path as black-holed and quarantine it
  // (100 ms, doubling up to 3.2 s); when the quarantine ends a few packets probe it again
  ps.loss_streak = 0;
  ps.banned_until_ns = now + (100000000ull << std::min<uint32_t>(ps.bans, 5));
  ++ps.bans;
  ++f.st.path_bans;
  UB_INFO(SUB_NET, "net: flow %u path %d quarantined (%u)", f.id, path, ps.bans);
}

void Engine::detect_loss(Flow& f, uint64_t now) {
  if (f.newest_acked_send_ts == 0) return;
  // RACK: a packet is lost once a packet sent sufficiently LATER has been acknowledged.  Time based, so
  // reordering between paths (which is the normal case here) does not trigger spurious retransmissions.
  const uint64_t reo_ns = (uint64_t)(std::min(std::max(f.srtt_us / 4 * f.reo_mult, 50.0), std::max(4 * f.srtt_us, 50.0)) * 1e3);
  const uint32_t n = f.snd_nxt - f.snd_una;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t s = f.snd_una + i;
    TxPkt& p = f.ring[s % kTxRing];
    if (!p.in_use || p.acked || p.lost || p.seq != s) continue;
    if (p.ts_send + reo_ns < f.newest_acked_send_ts) {
      p.lost = true;
      if (f.inflight) --f.inflight;
      if (f.path[p.path].inflight) --f.path[p.path].inflight;
      f.rexmit_q.push_back(s);
      ++f.st.fast_rexmit;
      ++est_.fast_rexmit;
      note_path_loss(f, p.path, now);
    }
  }
}

int Engine::pick_path(Flow& f, int avoid) {
  const int n = f.npaths;
  if (n <= 1) return 0;
  // power-of-two choices on (packets in flight, smoothed RTT) -- reference: select_qpidx_pot
  int a = (int)(rng_() % (uint64_t)n), b = (int)(rng_() % (uint64_t)n);
  if (a == avoid) a = (a + 1) % n;
  if (b == avoid) b = (b + 1) % n;
  const uint64_t now = now_ns();
  for (int tries = 0; tries < n && f.path[a].banned_until_ns > now; ++tries) a = (a + 1) % n;
  for (int tries = 0; tries < n && f.path[b].banned_until_ns > now; ++tries) b = (b + 1) % n;
  const PathState &pa = f.path[a], &pb = f.path[b];
  if (pa.inflight != pb.inflight) return pa.inflight < pb.inflight ? a : b;
  return pa.srtt_us <= pb.srtt_us ? a : b;
}

void Engine::emit_data(Flow& f, TxPkt& p, uint64_t now, bool is_rexmit) {
  const int path = pick_path(f, is_rexmit ? (int)p.path : -1);
  p.path = (uint16_t)path;
  p.ts_send = now;
  f.last_tx_ns = now;
  p.lost = false;
  if (is_rexmit) ++p.rexmits;
  PktHdr h{};
  h.magic = kMagic;
  h.type = PKT_DATA;
  h.kind = p.kind;
  h.path = p.path;
  h.dst_flow = f.peer_flow;
  h.seq = p.seq;
  h.ts_ns = now;
  h.msg_id = p.msg_id;
  h.len = p.len;
  h.offset = p.offset;
  h.msg_len = p.msg_len;
  if (cfg_.cc == CC_EQDS) {  // cumulative demand: everything sent so far + everything still queued
    uint64_t backlog = 0;
    for (size_t i = f.tx_cursor; i < f.txq.size(); ++i) backlog += f.txq[i]->len - f.txq[i]->next_off;
    h.aux = f.sent_payload_cum + backlog;
  } else {
    h.aux = 0;
  }
  ++f.inflight;
  ++f.path[path].inflight;
  ++f.path[path].tx;
  ++f.st.tx_pkts;
  ++f.st.path_tx[path];
  f.st.tx_bytes += p.len;
  raw_send(path, f.peer_addr[path], &h, sizeof(h), p.payload, p.len);
  if (cfg_.cc == CC_TIMELY) {
    const double gap_ns = (double)(p.len + sizeof(PktHdr) + 28) * 8.0 / f.timely.rate_gbps();
    const uint64_t floor = now > 100000 ? now - 100000 : 0;  // allow a 100 us burst
    f.pace_next_ns = std::max(f.pace_next_ns, floor) + (uint64_t)gap_ns;
  }
}

bool Engine::can_send_new(Flow& f, uint64_t now) {
  if ((uint32_t)(f.snd_nxt - f.snd_una) >= (uint32_t)(kSackBits - 1)) return false;
  uint32_t wnd = (uint32_t)cfg_.max_inflight;
  if (cfg_.cc == CC_SWIFT) wnd = std::min<uint32_t>(wnd, (uint32_t)std::max(1.0, f.swift.cwnd()));
  if (f.inflight >= wnd) return false;
  if (cfg_.cc == CC_TIMELY && now < f.pace_next_ns) return false;
  return true;
}

bool Engine::tx_pump(Flow& f, uint64_t now) {
  bool sent = false;
  // 1. retransmissions first (they do not wait for window space: the lost packet already left the window)
  while (!f.rexmit_q.empty()) {
    const uint32_t s = f.rexmit_q.front();
    TxPkt& p = f.ring[s % kTxRing];
    if (!(p.in_use && p.seq == s && !p.acked && p.lost)) {
      f.rexmit_q.pop_front();
      continue;
    }
    if (cfg_.cc == CC_TIMELY && now < f.pace_next_ns) break;
    f.rexmit_q.pop_front();
    emit_data(f, p, now, true);
    sent = true;
  }
  auto new_pkt = [&](uint8_t kind) -> TxPkt& {
    const uint32_t s = f.snd_nxt++;
    TxPkt& p = f.ring[s % kTxRing];
    p = TxPkt();
    p.seq = s;
    p.kind = kind;
    p.in_use = true;
    return p;
  };
  // 2. control frames.  A sender that is parked behind the receiver's RTR with nothing in flight probes the
  // peer once a second (an RTR frame of its own is a harmless reliable packet): a dead peer then surfaces
  // through the retransmission limit instead of an unbounded wait.
  // (With EQDS a lost credit ACK parks the sender the same way; there the probe goes out after a few RTTs -- its
  // ACK carries the receiver's current grant.)
  const uint64_t probe_after = (cfg_.cc == CC_EQDS && f.credit_starved)
                                   ? (uint64_t)(std::max(4 * f.srtt_us, 1000.0) * 1e3)
                                   : 1000000000ull;
  if (f.tx_cursor < f.txq.size() && f.snd_una == f.snd_nxt && now - std::max(f.last_progress_ns, f.last_tx_ns) > probe_after) {
    f.rtr_pending = true;
    // a closing flow must not look alive because of its own probes (and their ACKs): its give-up timer keeps running
    if (f.state.load(std::memory_order_relaxed) != FL_CLOSING) f.last_progress_ns = now;
    else f.last_tx_ns = now;
  }
  f.credit_starved = false;
  if (f.rtr_pending && can_send_new(f, now)) {
    TxPkt& p = new_pkt(FR_RTR);
    p.msg_id = f.rx_posted;
    f.rtr_pending = false;
    emit_data(f, p, now, false);
    sent = true;
  }
  // 3. message data, strictly in message order
  while (f.tx_cursor < f.txq.size() && can_send_new(f, now)) {
    TxMsg* m = f.txq[f.tx_cursor];
    const int32_t ahead = seq_diff(m->id, f.peer_posted);  // < 0: the receive is already posted
    const bool eager = m->len <= cfg_.eager_max && ahead < cfg_.eager_ahead;
    if (ahead >= 0 && !eager) break;
    const size_t chunk = std::min<size_t>((size_t)cfg_.payload, m->len - m->next_off);
    if (cfg_.cc == CC_EQDS) {
      const uint64_t spec = (uint64_t)cfg_.payload * 16;  // speculative first window, then credits
      if (f.sent_payload_cum + chunk > f.credit_cum + spec) {
        f.credit_starved = true;  // see the probe above
        break;
      }
    }
    TxPkt& p = new_pkt(FR_MSG);
    p.msg_id = m->id;
    p.offset = m->next_off;
    p.len = (uint32_t)chunk;
    p.msg_len = m->len;
    p.payload = m->ptr + m->next_off;
    p.msg = m;
    m->next_off += chunk;
    ++m->pkts_out;
    if (m->next_off >= m->len) {
      m->all_queued = true;
      ++f.tx_cursor;
    }
    f.sent_payload_cum += chunk;
    emit_data(f, p, now, false);
    sent = true;
  }
  if (f.fin_pending && !f.fin_sent && f.tx_cursor >= f.txq.size() && can_send_new(f, now)) {
    TxPkt& p = new_pkt(FR_FIN);
    f.fin_sent = true;
    emit_data(f, p, now, false);
    sent = true;
  }
  return sent;
}

// ------------------------------------------------------------------------------------------ timers
void Engine::timers(uint64_t now) {
  bool erased = false;
  for (auto& sp : active_) {
    Flow& f = *sp;
    const int st = f.state.load(std::memory_order_relaxed);
    if (st == FL_SYN_SENT) {
      if (now > f.syn_deadline_ns) {
        fail_flow(f, "connect timed out");
      } else if (now >= f.syn_next_ns) {
        send_syn(f, false, 0, nullptr);
        f.syn_next_ns = now + (uint64_t)cfg_.syn_retry_ms * 1000000ull;
      }
      continue;
    }
    if (st == FL_CLOSED) {
      if (now - f.last_progress_ns > kLingerNs) erased = true;
      continue;
    }
    if (st == FL_ERROR) {  // kept for a while so that the owner can still read the state, then reclaimed
      if (now - f.last_progress_ns > 15 * kLingerNs) erased = true;
      continue;
    }
    if (st != FL_ESTABLISHED && st != FL_CLOSING) continue;
    if (f.snd_una != f.snd_nxt) {
      TxPkt& p = f.ring[f.snd_una % kTxRing];
      const uint64_t rto = f.rto_ns << std::min(f.rto_count, 6);
      const uint64_t eff = std::min<uint64_t>(rto, (uint64_t)cfg_.rto_max_us * 1000ull);
      if (p.in_use && !p.acked && now > p.ts_send + eff) {
        ++f.rto_count;
        ++f.st.rto_rexmit;
        ++est_.rto_rexmit;
        if (f.rto_count >= cfg_.rto_abort) {
          fail_flow(f, "peer unreachable (retransmission limit)");
          continue;
        }
        if (!p.lost) {
          p.lost = true;
          if (f.inflight) --f.inflight;
          if (f.path[p.path].inflight) --f.path[p.path].inflight;
          note_path_loss(f, p.path, now);
        }
        p.ts_send = now;  // re-arm; the retransmission below stamps it again
        f.rexmit_q.push_front(f.snd_una);
        if (f.rto_count >= 3 && cfg_.cc == CC_SWIFT) f.swift.on_retransmit_timeout();
      } else if (!f.tlp_fired && f.rexmit_q.empty() && f.srtt_us > 0 &&
                 now > std::max(f.last_tx_ns, f.last_progress_ns) + (uint64_t)(std::max(2.0 * f.srtt_us, 300.0) * 1e3)) {
        // Tail-loss probe: the flight has been silent for two RTTs.  If the LAST packets of a burst were lost no
        // later packet will ever be acknowledged, so RACK cannot see the hole and only the (much longer) RTO
        // would; re-sending the newest unacknowledged packet draws an ACK whose SACK exposes the loss.
        for (uint32_t s = f.snd_nxt - 1; seq_diff(s, f.snd_una) >= 0; --s) {
          TxPkt& q = f.ring[s % kTxRing];
          if (q.in_use && q.seq == s && !q.acked) {
            if (!q.lost) {
              q.lost = true;
              if (f.inflight) --f.inflight;
              if (f.path[q.path].inflight) --f.path[q.path].inflight;
            }
            f.rexmit_q.push_back(s);
            ++f.st.tlp;
            break;
          }
          if (s == f.snd_una) break;
        }
        f.tlp_fired = true;
      }
    }
    if (st == FL_CLOSING) {
      const bool drained = f.fin_sent && f.snd_una == f.snd_nxt && f.txq.empty();
      // a peer that has already closed its side may be gone by now: do not wait long for it to acknowledge our FIN
      const uint64_t give_up = f.peer_fin.load(std::memory_order_relaxed) ? 100000000ull : kLingerNs;
      // hard bound: probes parked behind the peer's RTR are acknowledged and would refresh last_progress_ns forever
      const bool overdue = f.close_start_ns && now - f.close_start_ns > 4 * kLingerNs;
      if (drained || overdue || now - f.last_progress_ns > give_up) {
        for (auto& q : f.rxq)
          if (q.req) complete(q.req, 0, 3), q.req = nullptr;
        f.rxq.clear();
        if (!drained)  // gave up waiting for the peer: whatever was still queued will never be delivered
          for (TxMsg* m : f.txq)
            if (m->req) complete(m->req, 0, 1), m->req = nullptr;
        f.state.store(FL_CLOSED);
        f.last_progress_ns = now;
        if (cfg_.cc == CC_EQDS) pacer_.remove(f.id);
      }
    }
  }
  if (erased) {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto it = flows_.begin(); it != flows_.end();) {
      Flow& f = *it->second;
      const int fst = f.state.load();
      if ((fst == FL_CLOSED && now - f.last_progress_ns > kLingerNs) ||
          (fst == FL_ERROR && now - f.last_progress_ns > 15 * kLingerNs)) {
        for (auto s = syn_index_.begin(); s != syn_index_.end();)
          s = (s->second == f.id) ? syn_index_.erase(s) : std::next(s);
        for (TxMsg* m : f.txq) delete m;
        f.txq.clear();
        it = flows_.erase(it);
      } else {
        ++it;
      }
    }
    active_dirty_ = true;
  }
}

void Engine::fail_flow(Flow& f, const char* why) {
  if (why) UB_WARN("net: flow %u failed: %s", f.id, why);
  const int prev = f.state.exchange(FL_ERROR);
  if (prev == FL_ERROR) return;
  f.last_progress_ns = now_ns();
  for (TxMsg* m : f.txq) {
    if (m->req) complete(m->req, 0, 1);
    delete m;
  }
  f.txq.clear();
  f.tx_cursor = 0;
  for (auto& q : f.rxq)
    if (q.req) complete(q.req, 0, 1);
  f.rxq.clear();
  for (auto& p : f.ring) p = TxPkt();
  f.rexmit_q.clear();
  f.snd_una = f.snd_nxt;
  if (cfg_.cc == CC_EQDS) pacer_.remove(f.id);
  if (why && f.peer_flow && f.npaths > 0 && !stop_.load()) send_rst(0, f.peer_addr[0], f.peer_flow);
}

}  // namespace net
}  // namespace ub
