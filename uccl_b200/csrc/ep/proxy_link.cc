#include "proxy_link.h"

#include <string.h>

#include <chrono>

#include "../common/log.h"
#include "../common/timers.h"

namespace ub {

namespace {
constexpr uint32_t kLinkMagic = 0x4b4e4c50u;  // "PLNK"
enum : uint32_t { K_WRITE = 1, K_ADD = 2, K_NOTIFY = 3 };
}  // namespace

ProxyLink::ProxyLink(int box, int nboxes, std::shared_ptr<net::Engine> engine, std::vector<uint32_t> flows, WriteFn w,
                     AddFn a, NotifyFn n)
    : box_(box), n_(nboxes), eng_(std::move(engine)), write_(std::move(w)), add_(std::move(a)), notify_(std::move(n)) {
  UB_CHECK(nboxes >= 1 && box >= 0 && box < nboxes && (int)flows.size() == nboxes, "proxy link: bad box / flow table");
  peers_.assign((size_t)nboxes, nullptr);
  for (int k = 0; k < nboxes; ++k) {
    if (k == box) continue;
    peers_[k] = new Peer();
    peers_[k]->flow = flows[k];
  }
  rx_ = std::thread([this] { receiver(); });
}

ProxyLink::~ProxyLink() {
  try {
    flush(2000);
  } catch (...) {
  }
  stop_.store(true);
  if (rx_.joinable()) rx_.join();
  // posted header receives keep pointing into the Peer structs: they stay allocated (a few hundred bytes)
}

void ProxyLink::post(int dst_box, const Hdr& h, const void* payload) {
  UB_CHECK(dst_box >= 0 && dst_box < n_ && dst_box != box_, "proxy link: bad destination box %d", dst_box);
  auto p = std::make_unique<Pending>();
  p->hdr = h;
  if (h.kind == K_WRITE && h.bytes) p->payload.assign((const char*)payload, (const char*)payload + h.bytes);
  std::lock_guard<std::mutex> g(mu_);
  // header and payload enter the flow back to back under the lock: messages of different callers never interleave
  p->h = eng_->send_async(peers_[(size_t)dst_box]->flow, &p->hdr, sizeof(Hdr));
  if (!p->payload.empty()) p->p = eng_->send_async(peers_[(size_t)dst_box]->flow, p->payload.data(), p->payload.size());
  pending_.push_back(std::move(p));
  // reap what has completed so that the queue stays short
  while (!pending_.empty()) {
    Pending& f = *pending_.front();
    size_t nb = 0;
    int err = 0;
    if (f.h && eng_->test(f.h, &nb, &err)) f.h = nullptr;
    if (f.p && eng_->test(f.p, &nb, &err)) f.p = nullptr;
    if (f.h || f.p) break;
    pending_.pop_front();
  }
}

void ProxyLink::put(int dst_box, int dst_local, uint64_t dst_off, const void* src, uint32_t bytes) {
  Hdr h{kLinkMagic, K_WRITE, (uint32_t)dst_local, bytes, dst_off, 0};
  post(dst_box, h, src);
  std::lock_guard<std::mutex> g(mu_);
  ++st_.puts;
  st_.bytes_out += bytes;
}

void ProxyLink::add(int dst_box, int dst_local, uint64_t dst_off, uint64_t value) {
  Hdr h{kLinkMagic, K_ADD, (uint32_t)dst_local, 0, dst_off, value};
  post(dst_box, h, nullptr);
  std::lock_guard<std::mutex> g(mu_);
  ++st_.adds;
}

void ProxyLink::notify(int dst_box, uint32_t a, uint32_t b) {
  Hdr h{kLinkMagic, K_NOTIFY, 0, 0, a, b};
  post(dst_box, h, nullptr);
  std::lock_guard<std::mutex> g(mu_);
  ++st_.notifies;
}

void ProxyLink::flush(int timeout_ms) {
  const uint64_t t0 = now_ns();
  for (;;) {
    {
      std::lock_guard<std::mutex> g(mu_);
      while (!pending_.empty()) {
        Pending& f = *pending_.front();
        size_t nb = 0;
        int err = 0;
        if (f.h && eng_->test(f.h, &nb, &err)) f.h = nullptr;
        if (f.p && eng_->test(f.p, &nb, &err)) f.p = nullptr;
        if (f.h || f.p) break;
        pending_.pop_front();
      }
      if (pending_.empty()) return;
    }
    UB_CHECK(now_ns() - t0 < (uint64_t)timeout_ms * 1000000ull, "proxy link: flush timed out after %d ms", timeout_ms);
    std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}

ProxyLinkStats ProxyLink::stats() const {
  std::lock_guard<std::mutex> g(mu_);
  return st_;
}

void ProxyLink::receiver() {
  uint32_t idle = 0;
  while (!stop_.load(std::memory_order_relaxed)) {
    bool progress = false;
    for (int k = 0; k < n_; ++k) {
      Peer* s = peers_[k];
      if (!s || s->closed) continue;
      size_t nb = 0;
      int err = 0;
      if (!s->hdr_req && !s->pay_req) {
        s->hdr_req = eng_->recv_async(s->flow, &s->hdr, sizeof(Hdr));
        progress = true;
      }
      if (s->hdr_req && eng_->test(s->hdr_req, &nb, &err)) {
        s->hdr_req = nullptr;
        progress = true;
        if (err || nb != sizeof(Hdr) || s->hdr.magic != kLinkMagic) {
          if (err != 3 && !stop_.load()) UB_WARN("proxy link: bad frame from box %d (err %d)", k, err);
          s->closed = true;
          continue;
        }
        if (s->hdr.kind == K_WRITE && s->hdr.bytes) {
          s->buf.resize(s->hdr.bytes);
          s->pay_req = eng_->recv_async(s->flow, s->buf.data(), s->hdr.bytes);
        } else if (s->hdr.kind == K_ADD) {
          try {
            add_((int)s->hdr.dst_local, s->hdr.off, s->hdr.value);
          } catch (const std::exception& e) {  // a bad request of a peer must not take the process down
            UB_WARN("proxy link: ADD from box %d refused: %s", k, e.what());
            continue;
          }
          std::lock_guard<std::mutex> g(mu_);
          ++st_.applied_adds;
        } else if (s->hdr.kind == K_NOTIFY) {
          if (notify_) notify_(k, (uint32_t)s->hdr.off, (uint32_t)s->hdr.value);
          std::lock_guard<std::mutex> g(mu_);
          ++st_.applied_notifies;
        }
      }
      if (s->pay_req && eng_->test(s->pay_req, &nb, &err)) {
        s->pay_req = nullptr;
        progress = true;
        if (err || nb != s->hdr.bytes) {
          UB_WARN("proxy link: payload from box %d failed (err %d)", k, err);
          s->closed = true;
          continue;
        }
        try {
          write_((int)s->hdr.dst_local, s->hdr.off, s->buf.data(), s->hdr.bytes);
        } catch (const std::exception& e) {
          UB_WARN("proxy link: WRITE from box %d refused: %s", k, e.what());
          continue;
        }
        std::lock_guard<std::mutex> g(mu_);
        ++st_.applied_writes;
        st_.bytes_in += s->hdr.bytes;
      }
    }
    if (progress) idle = 0;
    else if (++idle > 200) std::this_thread::sleep_for(std::chrono::microseconds(20));
    else std::this_thread::yield();
  }
}

ProxyLink::WriteFn ProxyLink::host_write(std::vector<char*> heaps, uint64_t heap_bytes) {
  return [heaps, heap_bytes](int l, uint64_t off, const void* data, uint32_t bytes) {
    UB_CHECK(l >= 0 && l < (int)heaps.size() && off <= heap_bytes && bytes <= heap_bytes - off, "proxy link: WRITE outside the heap");
    memcpy(heaps[(size_t)l] + off, data, bytes);
  };
}

ProxyLink::AddFn ProxyLink::host_add(std::vector<char*> heaps, uint64_t heap_bytes) {
  return [heaps, heap_bytes](int l, uint64_t off, uint64_t value) {
    UB_CHECK(l >= 0 && l < (int)heaps.size() && off % 8 == 0 && off + 8 <= heap_bytes, "proxy link: ATOMIC outside the heap / unaligned");
    __atomic_fetch_add(reinterpret_cast<uint64_t*>(heaps[(size_t)l] + off), value, __ATOMIC_RELEASE);
  };
}

}  // namespace ub
