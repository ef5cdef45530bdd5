#include "uk_worker.h"

#include <sched.h>

#include <chrono>
#include <cstring>

#include "../coll/comm.h"
#include "../common/log.h"
#include "../common/param.h"
#include "../fabric/cu_api.h"

UB_PARAM(UkIdleUs, "UK_IDLE_US", 200000)

namespace ub {

namespace {
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
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

UkWorker::UkWorker(int device, int nlanes, uint64_t timeout_ms, int64_t idle_us) : device_(device), nlanes_(nlanes) {
  UB_CHECK(nlanes >= 1 && nlanes <= kUkMaxLanes, "ukernel: nlanes %d out of range (1..%d)", nlanes, kUkMaxLanes);
  timeout_ns_ = timeout_ms * 1000000ull;
  if (idle_us < 0) idle_us = ubParamUkIdleUs();
  idle_ns_ = (uint64_t)idle_us * 1000ull;
  lanes_.reset(new Lane[nlanes]);
  const size_t ring_bytes = sizeof(UkTask) * kUkRingEntries;
  if (is_host()) {
    for (int li = 0; li < nlanes_; ++li) {
      Lane& l = lanes_[li];
      void* p = nullptr;
      UB_CHECK(posix_memalign(&p, 64, ring_bytes) == 0, "ukernel: ring alloc failed");
      memset(p, 0, ring_bytes);
      l.ring = l.ring_dev = (UkTask*)p;
      UB_CHECK(posix_memalign(&p, 64, 64) == 0, "ukernel: counter alloc failed");
      memset(p, 0, 64);
      l.done_host = l.done_host_dev = l.done_dev = (uint64_t*)p;
    }
    err_host_ = err_dev_ = new uint32_t(0);
    return;
  }
  DevGuard g(device_);
  for (int li = 0; li < nlanes_; ++li) {
    Lane& l = lanes_[li];
    void* p = nullptr;
    UB_CUDA(cudaHostAlloc(&p, ring_bytes, cudaHostAllocMapped));
    memset(p, 0, ring_bytes);
    l.ring = (UkTask*)p;
    UB_CUDA(cudaHostGetDevicePointer((void**)&l.ring_dev, p, 0));
    UB_CUDA(cudaHostAlloc(&p, 64, cudaHostAllocMapped));
    memset(p, 0, 64);
    l.done_host = (uint64_t*)p;
    UB_CUDA(cudaHostGetDevicePointer((void**)&l.done_host_dev, p, 0));
    UB_CUDA(cudaMalloc((void**)&l.done_dev, 64));
  }
  void* e = nullptr;
  UB_CUDA(cudaHostAlloc(&e, 64, cudaHostAllocMapped));
  memset(e, 0, 64);
  err_host_ = (uint32_t*)e;
  UB_CUDA(cudaHostGetDevicePointer((void**)&err_dev_, e, 0));
  int lo = 0, hi = 0;
  UB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  UB_CUDA(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, hi));
  // no device-wide synchronisation here: in single-process worlds a peer rank's kernels may be
  // spinning on this very device waiting for us
  for (int li = 0; li < nlanes_; ++li) UB_CUDA(cudaMemsetAsync(lanes_[li].done_dev, 0, 64, stream_));
  UB_CUDA(cudaMalloc((void**)&votes_dev_, 64));
  UB_CUDA(cudaMemsetAsync(votes_dev_, 0, 64, stream_));
  UB_CUDA(cudaStreamSynchronize(stream_));
}

UkWorker::~UkWorker() {
  try {
    stop();
  } catch (...) {
  }
  if (is_host()) {
    for (int li = 0; li < nlanes_; ++li) {
      ::free(lanes_[li].ring);
      ::free(lanes_[li].done_host);
    }
    delete err_host_;
    return;
  }
  DevGuard g(device_);
  if (stream_) cudaStreamDestroy(stream_);
  for (int li = 0; li < nlanes_; ++li) {
    cudaFreeHost(lanes_[li].ring);
    cudaFreeHost(lanes_[li].done_host);
    cudaFree(lanes_[li].done_dev);
  }
  cudaFree(votes_dev_);
  cudaFreeHost(err_host_);
}

void UkWorker::start(cudaEvent_t wait_for) {
  if (running_) return;
  if (is_host()) {
    for (int i = 0; i < nlanes_; ++i) lanes_[i].start = lanes_[i].pushed.load();
    for (int i = 0; i < nlanes_; ++i) host_threads_.emplace_back([this, i] { host_lane_loop(i); });
    running_ = true;
    return;
  }
  DevGuard g(device_);
  // the kernel itself is launched by the first push (and relaunched whenever it has idled out);
  // every launch goes to stream_, i.e. after `wait_for`
  if (wait_for) UB_CUDA(cudaStreamWaitEvent(stream_, wait_for, 0));
  running_ = true;
  wd_stop_ = false;
  watchdog_ = std::thread([this] { watchdog_loop(); });
}

bool UkWorker::pending() const {
  for (int i = 0; i < nlanes_; ++i)
    if (pushed(i) > completed(i)) return true;
  return false;
}

void UkWorker::ensure_live_locked() {
  if (is_host() || !running_) return;
  if (kernel_live_) {
    cudaError_t q = cudaStreamQuery(stream_);
    if (q == cudaErrorNotReady) return;
    if (q != cudaSuccess) {
      UB_ERROR("ukernel: worker kernel failed: %s", cudaGetErrorString(q));
      __atomic_store_n(err_host_, 0x80000000u | 41u, __ATOMIC_RELEASE);
      return;
    }
    kernel_live_ = false;
    live_hint_ = false;
  }
  if (!pending()) return;
  UkWorkerArgs w;
  memset(&w, 0, sizeof(w));
  w.nlanes = nlanes_;
  w.err = err_dev_;
  w.timeout_ns = timeout_ns_;
  w.idle_ns = idle_ns_;
  w.votes = votes_dev_;
  for (int i = 0; i < nlanes_; ++i) {
    w.lane[i].ring = lanes_[i].ring_dev;
    w.lane[i].done_host = lanes_[i].done_host_dev;
    w.lane[i].done_dev = lanes_[i].done_dev;
    w.lane[i].start = completed(i);  // the previous instance stopped at a task boundary
  }
  cudaError_t e = cudaMemsetAsync(votes_dev_, 0, 4, stream_);
  if (e == cudaSuccess) e = launch_uk_worker(w, stream_);
  UB_CHECK(e == cudaSuccess, "ukernel: worker launch failed: %s", cudaGetErrorString(e));
  kernel_live_ = true;
  live_hint_ = true;
  ++launches_;
}

void UkWorker::kick() {
  if (is_host()) return;
  DevGuard g(device_);
  std::lock_guard<std::mutex> lk(launch_mu_);
  ensure_live_locked();
}

void UkWorker::watchdog_loop() {
  cudaSetDevice(device_);
  while (!wd_stop_.load(std::memory_order_acquire)) {
    std::this_thread::sleep_for(std::chrono::microseconds(500));
    try {
      std::lock_guard<std::mutex> lk(launch_mu_);
      ensure_live_locked();
    } catch (const std::exception& e) {
      UB_ERROR("ukernel watchdog: %s", e.what());
      return;
    }
  }
}

void UkWorker::stop() {
  if (!running_) return;
  UkTask t;
  memset(&t, 0, sizeof(t));
  t.op = UK_EXIT;
  if (is_host()) {
    for (int i = 0; i < nlanes_; ++i) push(i, t);
    for (auto& th : host_threads_) th.join();
    host_threads_.clear();
    running_ = false;
    return;
  }
  wd_stop_ = true;
  if (watchdog_.joinable()) watchdog_.join();
  DevGuard g(device_);
  if (error() == 0) {
    // drain outstanding work, then make a live kernel return right away instead of idling out
    for (int i = 0; i < nlanes_; ++i) push(i, t);
    kick();
  }
  cudaError_t e = cudaStreamSynchronize(stream_);
  if (e != cudaSuccess) UB_WARN("ukernel: worker exited with %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(launch_mu_);
  kernel_live_ = false;
  running_ = false;
}

uint64_t UkWorker::completed(int lane) const {
  return __atomic_load_n(lanes_[lane].done_host, __ATOMIC_ACQUIRE);
}

uint32_t UkWorker::error() const { return err_host_ ? __atomic_load_n(err_host_, __ATOMIC_ACQUIRE) : 0; }

uint64_t UkWorker::push(int lane, UkTask t) {
  UB_CHECK(lane >= 0 && lane < nlanes_, "ukernel: bad lane %d", lane);
  Lane& l = lanes_[lane];
  // ring full: wait for the consumer (bounded by the worker's own WAIT timeout)
  const uint64_t cur = l.pushed.load(std::memory_order_relaxed);
  if (cur - completed(lane) >= (uint64_t)kUkRingEntries) {
    const double t0 = now_s();
    kick();
    while (cur - completed(lane) >= (uint64_t)kUkRingEntries) {
      UB_CHECK(error() == 0, "ukernel: worker reported error 0x%x", error());
      UB_CHECK(running_, "ukernel: FIFO full and the worker is not running");
      UB_CHECK(now_s() - t0 < 60.0, "ukernel: FIFO of lane %d stuck for 60 s", lane);
      sched_yield();
      if ((++push_tick_ & 0xfff) == 0) kick();
    }
  }
  UkTask* slot = l.ring + (cur & (kUkRingEntries - 1));
  const uint64_t seq = cur + 1;
  t.seq = 0;
  // payload first (everything but the sequence word), then release the sequence word
  memcpy(slot, &t, offsetof(UkTask, seq));
  __atomic_store_n(&slot->seq, seq, __ATOMIC_RELEASE);
  l.pushed.store(seq, std::memory_order_release);
  // the kernel may have idled out (or was never launched): the first push after that relaunches it
  // here, later ones leave the check to the watchdog (cudaStreamQuery per task would dominate)
  if (!is_host() && (!live_hint_.load(std::memory_order_relaxed) || ((++push_tick_) & 0xff) == 0)) kick();
  ++stats_.pushed;
  switch (t.op) {
    case UK_COPY: ++stats_.copies; stats_.bytes += t.bytes; break;
    case UK_REDUCE: ++stats_.reduces; stats_.bytes += t.bytes; break;
    case UK_SIGNAL: ++stats_.signals; break;
    case UK_WAIT: ++stats_.waits; break;
    default: break;
  }
  return seq;
}

void UkWorker::wait(int lane, uint64_t ticket, double timeout_s) const {
  const double t0 = now_s();
  uint32_t spins = 0;
  while (!done(lane, ticket)) {
    if ((++spins & 0xff) == 0) {
      UB_CHECK(error() == 0, "ukernel: worker reported error 0x%x", error());
      UB_CHECK(now_s() - t0 < timeout_s, "ukernel: lane %d ticket %llu not complete after %.1f s (completed %llu)", lane,
               (unsigned long long)ticket, timeout_s, (unsigned long long)completed(lane));
      sched_yield();
      if ((spins & 0xffff) == 0) const_cast<UkWorker*>(this)->kick();
    }
  }
  // a lane that gave up on a WAIT publishes "done" to unblock its waiters: completion alone is not success
  UB_CHECK(error() == 0, "ukernel: worker reported error 0x%x", error());
}

void UkWorker::wait_all(double timeout_s) const {
  for (int i = 0; i < nlanes_; ++i) wait(i, pushed(i), timeout_s);
}

// ---------------------------------------------------------------- host interpreter
void UkWorker::host_lane_loop(int lane) {
  Lane& l = lanes_[lane];
  for (uint64_t seq = l.start;; ++seq) {
    UkTask* slot = l.ring + (seq & (kUkRingEntries - 1));
    while (__atomic_load_n(&slot->seq, __ATOMIC_ACQUIRE) != seq + 1) sched_yield();
    UkTask t;
    memcpy(&t, slot, sizeof(t));
    switch (t.op) {
      case UK_COPY: memmove((void*)t.dst, (const void*)t.src, t.bytes); break;
      case UK_REDUCE: {
        const void* srcs[2] = {(const void*)t.src, (const void*)t.src2};
        host_reduce_n((void*)t.dst, srcs, 2, t.bytes / dtype_size((int)t.dtype), (int)t.dtype, (int)t.redop, 1.0f);
        break;
      }
      case UK_SIGNAL: __atomic_fetch_add((uint64_t*)t.sig_addr, t.sig_val, __ATOMIC_ACQ_REL); break;
      case UK_WAIT: {
        const double t0 = now_s();
        uint32_t spins = 0;
        while (__atomic_load_n((uint64_t*)t.sig_addr, __ATOMIC_ACQUIRE) < t.sig_val) {
          if ((++spins & 0xff) == 0) {
            if (timeout_ns_ && (now_s() - t0) * 1e9 > (double)timeout_ns_) {
              UB_ERROR("ukernel(host): lane %d task %llu WAIT timeout", lane, (unsigned long long)seq);
              __atomic_store_n(err_host_, 0x80000000u | 40u, __ATOMIC_RELEASE);
              __atomic_store_n(l.done_host, seq + 1, __ATOMIC_RELEASE);
              return;
            }
            sched_yield();
          }
        }
        break;
      }
      default: break;
    }
    __atomic_store_n(l.done_host, seq + 1, __ATOMIC_RELEASE);
    if (t.op == UK_EXIT) return;
  }
}

}  // namespace ub
