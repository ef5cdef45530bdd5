#include "proxy.h"

#include <sched.h>

#include <chrono>
#include <cstring>

#include "../common/log.h"
#include "../fabric/cu_api.h"

namespace ub {

namespace {
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct DevGuard {
  int prev = -1;
  bool active = false;
  explicit DevGuard(int dev) {
    if (dev < 0) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      cudaSetDevice(dev);
      active = true;
    }
  }
  ~DevGuard() {
    if (active) cudaSetDevice(prev);
  }
};
}  // namespace

Proxy::Proxy(std::shared_ptr<Comm> comm, uint32_t capacity) : comm_(comm), capacity_(capacity) {
  UB_CHECK(!comm_->is_host(), "Proxy needs a CUDA communicator");
  UB_CHECK(capacity >= 64 && (capacity & (capacity - 1)) == 0, "proxy: capacity must be a power of two >= 64");
  DevGuard g(comm_->device());
  void* p = nullptr;
  UB_CUDA(cudaHostAlloc(&p, sizeof(D2HCmd) * capacity, cudaHostAllocMapped));
  memset(p, 0, sizeof(D2HCmd) * capacity);
  ring_ = (D2HCmd*)p;
  UB_CUDA(cudaHostGetDevicePointer((void**)&dev_.ring, p, 0));
  UB_CUDA(cudaHostAlloc(&p, 128, cudaHostAllocMapped));
  memset(p, 0, 128);
  tail_ = (uint64_t*)p;
  ack_ = tail_ + 8;  // separate cache line
  void* d = nullptr;
  UB_CUDA(cudaHostGetDevicePointer(&d, p, 0));
  dev_.tail = (const volatile uint64_t*)d;
  dev_.ack = (volatile uint64_t*)d + 8;
  UB_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  UB_CUDA(cudaMalloc((void**)&dev_.head, 64));
  UB_CUDA(cudaMemsetAsync(dev_.head, 0, 64, stream_));
  UB_CUDA(cudaStreamSynchronize(stream_));
  dev_.capacity = capacity;
}

Proxy::~Proxy() {
  try {
    stop();
  } catch (...) {
  }
  link_.reset();  // stops the link's receiver before its stream goes away
  DevGuard g(comm_->device());
  if (link_stream_) {
    cudaStreamSynchronize(link_stream_);
    cudaStreamDestroy(link_stream_);
  }
  if (bounce_) cudaFreeHost(bounce_);
  if (stream_) cudaStreamDestroy(stream_);
  if (dev_.head) cudaFree(dev_.head);
  if (ring_) cudaFreeHost(ring_);
  if (tail_) cudaFreeHost(tail_);
}

void Proxy::start() {
  if (running_) return;
  stop_ = false;
  th_ = std::thread([this] { loop(); });
  running_ = true;
}

void Proxy::stop() {
  if (!running_) return;
  stop_ = true;
  th_.join();
  running_ = false;
}

void Proxy::loop() {
  cudaSetDevice(comm_->device());
  uint64_t idx = __atomic_load_n(tail_, __ATOMIC_ACQUIRE);
  uint32_t idle = 0;
  while (!stop_.load(std::memory_order_acquire)) {
    D2HCmd* slot = ring_ + (idx & (capacity_ - 1));
    const uint32_t want = (uint32_t)(idx + 1) | 0x80000000u;
    if (__atomic_load_n(&slot->tag, __ATOMIC_ACQUIRE) != want) {
      if (++idle > 2000) {
        sched_yield();
        idle = 0;
      }
      continue;
    }
    idle = 0;
    D2HCmd c;
    memcpy(&c, slot, sizeof(c));
    const double t0 = now_s();
    const uint32_t type = c.type_dst_aux & 0xffu;
    try {
      handle(c);
    } catch (const std::exception& e) {
      UB_ERROR("proxy: command %llu (type %u) failed: %s", (unsigned long long)idx, type, e.what());
    }
    {
      std::lock_guard<std::mutex> g(mu_);
      ++stats_.cmds;
      handle_us_sum_ += (now_s() - t0) * 1e6;
    }
    ++idx;
    __atomic_store_n(tail_, idx, __ATOMIC_RELEASE);
    if (type == D2H_QUIT) break;
  }
}

void Proxy::attach_link(std::shared_ptr<net::Engine> engine, std::vector<uint32_t> flows, int box, int nboxes,
                        int local_size) {
  UB_CHECK(!running_, "proxy: attach_link before start()");
  UB_CHECK(local_size == comm_->nranks(), "proxy: local_size (%d) must equal the box communicator's size (%d)", local_size,
           comm_->nranks());
  box_ = box;
  local_size_ = local_size;
  DevGuard g(comm_->device());
  UB_CUDA(cudaStreamCreateWithFlags(&link_stream_, cudaStreamNonBlocking));
  const int dev = comm_->device();
  auto check = [this](int l, uint64_t off, uint64_t bytes) {
    UB_CHECK(l >= 0 && l < comm_->nranks() && off + bytes <= comm_->fabric().heap_bytes(), "proxy link: access outside the heap");
  };
  // inbound: applied on one stream, so an ATOMIC lands after every WRITE that preceded it on the wire
  ProxyLink::WriteFn w = [this, dev, check](int l, uint64_t off, const void* data, uint32_t bytes) {
    check(l, off, bytes);
    cudaSetDevice(dev);
    // pageable source: cudaMemcpyAsync returns once the data is staged, the link may reuse its buffer
    UB_CUDA(cudaMemcpyAsync(comm_->fabric().heap(l) + off, data, bytes, cudaMemcpyHostToDevice, link_stream_));
  };
  ProxyLink::AddFn a = [this, dev, check](int l, uint64_t off, uint64_t value) {
    check(l, off, 8);
    cudaSetDevice(dev);
    cudaError_t e = launch_u64_add((uint64_t*)(comm_->fabric().heap(l) + off), value, link_stream_);
    UB_CHECK(e == cudaSuccess, "proxy link: atomic launch failed: %s", cudaGetErrorString(e));
  };
  ProxyLink::NotifyFn nf = [this](int src_box, uint32_t a_, uint32_t b_) {
    std::lock_guard<std::mutex> lk(mu_);
    notifs_.emplace_back(a_ | ((uint32_t)src_box << 24), b_);
  };
  link_.reset(new ProxyLink(box, nboxes, std::move(engine), std::move(flows), std::move(w), std::move(a), std::move(nf)));
}

void Proxy::handle(const D2HCmd& c) {
  const uint32_t type = c.type_dst_aux & 0xffu, aux = c.type_dst_aux >> 16;
  uint32_t dst = (c.type_dst_aux >> 8) & 0xffu;
  if (link_) {  // destinations are global ranks (box major): split into (box, local rank)
    const int dbox = (int)dst / local_size_;
    const int dl = (int)dst % local_size_;
    if (dbox != box_ && (type == D2H_WRITE || type == D2H_ATOMIC)) {
      UB_CHECK(dbox < link_->nboxes(), "proxy: bad destination rank %u", dst);
      if (type == D2H_WRITE) {
        const Fabric& f = comm_->fabric();
        UB_CHECK(c.src_off + c.bytes <= f.heap_bytes(), "proxy: WRITE source outside the heap");
        if (bounce_cap_ < c.bytes) {
          if (bounce_) cudaFreeHost(bounce_);
          bounce_cap_ = std::max<size_t>(c.bytes, 1 << 20);
          UB_CUDA(cudaHostAlloc((void**)&bounce_, bounce_cap_, cudaHostAllocDefault));
        }
        UB_CUDA(cudaMemcpyAsync(bounce_, f.heap(comm_->rank()) + c.src_off, c.bytes, cudaMemcpyDeviceToHost, stream_));
        UB_CUDA(cudaStreamSynchronize(stream_));
        link_->put(dbox, dl, c.dst_off, bounce_, c.bytes);  // copies the bounce buffer
        std::lock_guard<std::mutex> g(mu_);
        ++stats_.writes;
        stats_.bytes += c.bytes;
      } else {
        link_->add(dbox, dl, c.dst_off, c.value);  // same flow as the puts: ordered after them
        std::lock_guard<std::mutex> g(mu_);
        ++stats_.atomics;
      }
      return;
    }
    dst = (uint32_t)dl;
  }
  switch (type) {
    case D2H_NOP: {
      __atomic_fetch_add(ack_, 1, __ATOMIC_RELEASE);
      std::lock_guard<std::mutex> g(mu_);
      ++stats_.nops;
      break;
    }
    case D2H_WRITE: {
      UB_CHECK((int)dst < comm_->nranks(), "proxy: bad destination rank %u", dst);
      const Fabric& f = comm_->fabric();
      UB_CHECK(c.src_off + c.bytes <= f.heap_bytes() && c.dst_off + c.bytes <= f.heap_bytes(),
               "proxy: WRITE outside the heap");
      UB_CUDA(cudaMemcpyAsync(f.heap((int)dst) + c.dst_off, f.heap(comm_->rank()) + c.src_off, c.bytes,
                              cudaMemcpyDeviceToDevice, stream_));
      std::lock_guard<std::mutex> g(mu_);
      ++stats_.writes;
      stats_.bytes += c.bytes;
      break;
    }
    case D2H_ATOMIC: {
      UB_CHECK((int)dst < comm_->nranks(), "proxy: bad destination rank %u", dst);
      const Fabric& f = comm_->fabric();
      UB_CHECK(c.dst_off + 8 <= f.heap_bytes() && c.dst_off % 8 == 0, "proxy: ATOMIC outside the heap / unaligned");
      // same stream as the copies: the counter moves only after every earlier WRITE has landed
      cudaError_t e = launch_u64_add((uint64_t*)(f.heap((int)dst) + c.dst_off), c.value, stream_);
      UB_CHECK(e == cudaSuccess, "proxy: atomic launch failed: %s", cudaGetErrorString(e));
      std::lock_guard<std::mutex> g(mu_);
      ++stats_.atomics;
      break;
    }
    case D2H_NOTIFY: {
      std::lock_guard<std::mutex> g(mu_);
      notifs_.emplace_back(aux, c.value);
      ++stats_.notifies;
      break;
    }
    default: break;
  }
}

void Proxy::drain(double timeout_s) {
  DevGuard g(comm_->device());
  unsigned long long head = 0;
  // the head lives in device memory; read it through a side stream so running kernels are not waited for
  cudaStream_t s;
  UB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  const double t0 = now_s();
  while (true) {
    UB_CUDA(cudaMemcpyAsync(&head, dev_.head, sizeof(head), cudaMemcpyDeviceToHost, s));
    UB_CUDA(cudaStreamSynchronize(s));
    if (consumed() >= head) break;
    UB_CHECK(now_s() - t0 < timeout_s, "proxy: drain timed out (consumed %llu of %llu)", (unsigned long long)consumed(),
             head);
    sched_yield();
  }
  cudaStreamDestroy(s);
  UB_CUDA(cudaStreamSynchronize(stream_));
}

std::vector<std::pair<uint32_t, uint32_t>> Proxy::poll_notifications() {
  std::lock_guard<std::mutex> g(mu_);
  std::vector<std::pair<uint32_t, uint32_t>> out(notifs_.begin(), notifs_.end());
  notifs_.clear();
  return out;
}

ProxyStats Proxy::stats() const {
  std::lock_guard<std::mutex> g(mu_);
  ProxyStats s = stats_;
  s.avg_handle_us = s.cmds ? handle_us_sum_ / (double)s.cmds : 0.0;
  return s;
}

double Proxy::bench_throughput(int blocks, int threads, int per_thread, cudaStream_t st) {
  UB_CHECK(running_, "proxy not running");
  DevGuard g(comm_->device());
  const uint64_t before = consumed();
  const uint64_t total = (uint64_t)blocks * threads * per_thread;
  const double t0 = now_s();
  cudaError_t e = launch_d2h_bench(dev_, blocks, threads, per_thread, st);
  UB_CHECK(e == cudaSuccess, "bench launch failed: %s", cudaGetErrorString(e));
  while (consumed() < before + total) {
    UB_CHECK(now_s() - t0 < 60.0, "proxy bench timed out");
    sched_yield();
  }
  const double dt = now_s() - t0;
  UB_CUDA(cudaStreamSynchronize(st));
  return (double)total / dt;
}

double Proxy::bench_latency(int iters, cudaStream_t st) {
  UB_CHECK(running_, "proxy not running");
  DevGuard g(comm_->device());
  unsigned long long* d = nullptr;
  UB_CUDA(cudaMalloc((void**)&d, 8));
  cudaError_t e = launch_d2h_latency(dev_, iters, d, st);
  UB_CHECK(e == cudaSuccess, "latency launch failed: %s", cudaGetErrorString(e));
  unsigned long long total = 0;
  UB_CUDA(cudaMemcpyAsync(&total, d, 8, cudaMemcpyDeviceToHost, st));
  UB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d);
  return (double)total / 1e3 / (double)iters;
}

void Proxy::issue_from_device(uint32_t type, int dst_rank, uint32_t aux, uint64_t src_off, uint64_t dst_off,
                              uint32_t bytes, uint32_t value, cudaStream_t st) {
  DevGuard g(comm_->device());
  cudaError_t e = launch_d2h_issue(dev_, type, (uint32_t)dst_rank, aux, src_off, dst_off, bytes, value, st);
  UB_CHECK(e == cudaSuccess, "issue launch failed: %s", cudaGetErrorString(e));
}

}  // namespace ub
