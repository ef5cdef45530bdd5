#include "uk_comm.h"

#include <algorithm>
#include <cstring>

#include "../common/log.h"
#include "../fabric/cu_api.h"

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
uint64_t round_up(uint64_t a, uint64_t b) { return (a + b - 1) / b * b; }
}  // namespace

UkComm::UkComm(std::shared_ptr<Comm> comm, const UkCommConfig& cfg) : comm_(comm), cfg_(cfg) {
  const int n = comm_->nranks(), L = cfg_.nlanes;
  UB_CHECK(L >= 1 && L <= kUkMaxLanes, "ukernel: nlanes %d out of range", L);
  UB_CHECK(cfg_.tile_bytes >= 16 && cfg_.tile_bytes % 16 == 0, "ukernel: tile_bytes must be a multiple of 16");
  cfg_.staging_bytes = round_up(std::max<uint64_t>(cfg_.staging_bytes, 16ull * n), 16ull * n);
  const uint64_t ctrl_bytes = round_up(((uint64_t)L * n + 2) * sizeof(uint64_t), 256);
  const uint64_t scratch_bytes = std::max(uk_scratch_bytes(UkAlgo::Ring, n, L, cfg_.tile_bytes),
                                          uk_scratch_bytes(UkAlgo::FullMesh, n, L, cfg_.tile_bytes));
  // symmetric allocations: every rank performs the same sequence, so offsets agree (same
  // contract as EpBuffer); peers are addressed with Comm::peer_ptr
  ctrl_ = (char*)comm_->alloc(ctrl_bytes, 256);
  scratch_ = scratch_bytes ? (char*)comm_->alloc(scratch_bytes, 256) : nullptr;
  stage_in_ = (char*)comm_->alloc(cfg_.staging_bytes, 256);
  stage_out_ = (char*)comm_->alloc(cfg_.staging_bytes, 256);
  flags_ = (uint64_t*)ctrl_;
  lane_sync_ = flags_ + (uint64_t)L * n;
  ready_ = lane_sync_ + 1;
  expected_.assign((size_t)L * n, 0);
  my_offs_[0] = comm_->heap_offset(ctrl_);
  my_offs_[1] = scratch_ ? comm_->heap_offset(scratch_) : 0;
  my_offs_[2] = comm_->heap_offset(stage_in_);
  my_offs_[3] = comm_->heap_offset(stage_out_);
  region_bytes_[0] = ctrl_bytes, region_bytes_[1] = scratch_bytes;
  region_bytes_[2] = region_bytes_[3] = cfg_.staging_bytes;
  all_offs_.assign((size_t)n * kRegions, 0);
  worker_.reset(new UkWorker(comm_->is_host() ? -1 : comm_->device(), L));
  if (comm_->is_host()) {
    memset(ctrl_, 0, ctrl_bytes);
    comm_->allgather(my_offs_, all_offs_.data(), kRegions, kU64, nullptr);
    offs_ready_ = true;
    comm_->barrier(nullptr);  // every rank has zeroed its counters before anybody signals
    worker_->start();
    return;
  }
  DevGuard g(comm_->device());
  UB_CUDA(cudaStreamCreateWithFlags(&setup_stream_, cudaStreamNonBlocking));
  UB_CUDA(cudaMemsetAsync(ctrl_, 0, ctrl_bytes, setup_stream_));
  // all-gather the region offsets on the device; the table is read back lazily (first operation),
  // because this constructor must not block on peers that are constructed later by the same thread
  UB_CUDA(cudaMalloc((void**)&offs_dev_, sizeof(uint64_t) * kRegions * (n + 1)));
  UB_CUDA(cudaHostAlloc((void**)&offs_host_, sizeof(uint64_t) * kRegions * n, cudaHostAllocDefault));
  UB_CUDA(cudaMemcpyAsync(offs_dev_, my_offs_, sizeof(my_offs_), cudaMemcpyHostToDevice, setup_stream_));
  comm_->allgather(offs_dev_, offs_dev_ + kRegions, kRegions, kU64, setup_stream_);
  UB_CUDA(cudaMemcpyAsync(offs_host_, offs_dev_ + kRegions, sizeof(uint64_t) * kRegions * n, cudaMemcpyDeviceToHost,
                          setup_stream_));
  // the cross-rank barrier runs on the device; the worker kernel is ordered after it, so this
  // constructor never blocks on a peer (ranks of a single-process world are built one by one)
  comm_->barrier(setup_stream_);
  cudaEvent_t ev;
  UB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  UB_CUDA(cudaEventRecord(ev, setup_stream_));
  worker_->start(ev);
  UB_CUDA(cudaEventDestroy(ev));
  stream_ops_ = cu().StreamWriteValue64 != nullptr && cu().StreamWaitValue64 != nullptr;
}

UkComm::~UkComm() {
  try {
    stop();
  } catch (...) {
  }
  if (setup_stream_) {
    DevGuard g(comm_->device());
    cudaStreamSynchronize(setup_stream_);
    cudaStreamDestroy(setup_stream_);
    if (offs_dev_) cudaFree(offs_dev_);
    if (offs_host_) cudaFreeHost(offs_host_);
  }
  worker_.reset();
  if (stage_out_) comm_->free(stage_out_);
  if (stage_in_) comm_->free(stage_in_);
  if (scratch_) comm_->free(scratch_);
  if (ctrl_) comm_->free(ctrl_);
}

void UkComm::stop() {
  if (worker_ && worker_->running()) worker_->stop();
}

void UkComm::push(int lane, const UkTask& t) {
  worker_->push(lane, t);
  ++stats_.tasks;
}

// every lane arrives, then waits for all lanes: orders lane-sliced copies against plan ops that
// read / write across slices
void UkComm::lane_barrier() {
  const int L = cfg_.nlanes;
  ++lane_barriers_;
  for (int l = 0; l < L; ++l) {
    UkTask s;
    memset(&s, 0, sizeof(s));
    s.op = UK_SIGNAL;
    s.sig_addr = (uint64_t)lane_sync_;
    s.sig_val = 1;
    push(l, s);
    UkTask w;
    memset(&w, 0, sizeof(w));
    w.op = UK_WAIT;
    w.sig_addr = (uint64_t)lane_sync_;
    w.sig_val = lane_barriers_ * (uint64_t)L;
    push(l, w);
  }
}

void UkComm::resolve_offsets() {
  if (offs_ready_) return;
  DevGuard g(comm_->device());
  UB_CUDA(cudaStreamSynchronize(setup_stream_));  // every rank has been constructed by the time of the first op
  memcpy(all_offs_.data(), offs_host_, sizeof(uint64_t) * all_offs_.size());
  offs_ready_ = true;
}

// Address, in this rank's address space, of peer `peer`'s instance of the buffer `local_ptr` points into.
char* UkComm::remote(const char* local_ptr, int peer) const {
  const char* bases[kRegions] = {ctrl_, scratch_, stage_in_, stage_out_};
  for (int r = 0; r < kRegions; ++r) {
    if (bases[r] && local_ptr >= bases[r] && local_ptr < bases[r] + region_bytes_[r])
      return comm_->fabric().heap(peer) + all_offs_[(size_t)peer * kRegions + r] + (local_ptr - bases[r]);
  }
  // a caller-declared symmetric user buffer: same offset everywhere
  return (char*)comm_->peer_ptr(local_ptr, peer);
}

void UkComm::begin_op(cudaStream_t stream) {
  UB_CHECK(worker_->running(), "ukernel: communicator was stopped");
  resolve_offsets();
  UB_CHECK(worker_->error() == 0, "ukernel: worker reported error 0x%x", worker_->error());
  ++ops_;
  ++stats_.ops;
  if (comm_->is_host()) return;
  DevGuard g(comm_->device());
  bool ordered = false;
  if (stream_ops_) {
    // "everything enqueued on `stream` so far is done" -> ready counter; each lane waits for it
    CUresult r = cu().StreamWriteValue64((CUstream)stream, (CUdeviceptr)ready_, ops_, 0);
    if (r == CUDA_SUCCESS) {
      for (int l = 0; l < cfg_.nlanes; ++l) {
        UkTask w;
        memset(&w, 0, sizeof(w));
        w.op = UK_WAIT;
        w.sig_addr = (uint64_t)ready_;
        w.sig_val = ops_;
        push(l, w);
      }
      ordered = true;
      ++stats_.stream_ordered_ops;
    } else {
      UB_WARN("ukernel: cuStreamWriteValue64 failed (%s); falling back to host synchronisation", cu_errstr(r));
      stream_ops_ = false;
    }
  }
  if (!ordered) UB_CUDA(cudaStreamSynchronize(stream));
}

uint64_t UkComm::end_op(cudaStream_t stream) {
  lane_barrier();  // all lanes finished; lane_sync_ == lane_barriers_ * L marks completion of the op
  std::vector<uint64_t> t(cfg_.nlanes);
  for (int l = 0; l < cfg_.nlanes; ++l) t[l] = worker_->pushed(l);
  tickets_[ops_] = t;
  while (tickets_.size() > 4096) tickets_.erase(tickets_.begin());
  if (!comm_->is_host() && stream_ops_) {
    DevGuard g(comm_->device());
    CUresult r = cu().StreamWaitValue64((CUstream)stream, (CUdeviceptr)lane_sync_, lane_barriers_ * (uint64_t)cfg_.nlanes,
                                        CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) {
      UB_WARN("ukernel: cuStreamWaitValue64 failed (%s); callers must wait() on the host", cu_errstr(r));
      stream_ops_ = false;
    }
  }
  return ops_;
}

bool UkComm::test(uint64_t ticket) {
  auto it = tickets_.find(ticket);
  if (it == tickets_.end()) return true;  // long gone
  for (int l = 0; l < cfg_.nlanes; ++l)
    if (!worker_->done(l, it->second[l])) return false;
  return true;
}

void UkComm::wait(uint64_t ticket, double timeout_s) {
  auto it = tickets_.find(ticket);
  if (it == tickets_.end()) return;
  for (int l = 0; l < cfg_.nlanes; ++l) worker_->wait(l, it->second[l], timeout_s);
}

char* UkComm::local(const Bufs& b, const UkRef& r) const {
  switch (r.buf) {
    case UkBuf::In: return b.in + r.off;
    case UkBuf::Out: return b.out + r.off;
    default: return scratch_ + r.off;
  }
}

void UkComm::run_plan(const UkPlan& plan, const Bufs& b, int dtype, int op) {
  const int n = comm_->nranks();
  for (const UkPlanOp& o : plan.ops) {
    UkTask t;
    memset(&t, 0, sizeof(t));
    switch (o.kind) {
      case UkPlanOp::Copy:
        if (local(b, o.dst) == local(b, o.src)) break;
        t.op = UK_COPY;
        t.dst = (uint64_t)local(b, o.dst);
        t.src = (uint64_t)local(b, o.src);
        t.bytes = o.bytes;
        push(o.lane, t);
        break;
      case UkPlanOp::Reduce:
        t.op = UK_REDUCE;
        t.dtype = (uint32_t)dtype;
        t.redop = (uint32_t)(op == kAvg ? kSum : op);
        t.dst = (uint64_t)local(b, o.dst);
        t.src = (uint64_t)local(b, o.src);
        t.src2 = (uint64_t)local(b, o.src2);
        t.bytes = o.bytes;
        push(o.lane, t);
        break;
      case UkPlanOp::Send: {
        if (o.bytes) {
          t.op = UK_COPY;
          t.dst = (uint64_t)remote(local(b, o.dst), o.peer);
          t.src = (uint64_t)local(b, o.src);
          t.bytes = o.bytes;
          push(o.lane, t);
        }
        UkTask s;
        memset(&s, 0, sizeof(s));
        s.op = UK_SIGNAL;
        s.sig_addr = (uint64_t)remote((const char*)(flags_ + (uint64_t)o.lane * n + comm_->rank()), o.peer);
        s.sig_val = 1;
        push(o.lane, s);
        break;
      }
      case UkPlanOp::Recv:
        t.op = UK_WAIT;
        t.sig_addr = (uint64_t)(flags_ + (uint64_t)o.lane * n + o.peer);
        t.sig_val = ++expected_[(size_t)o.lane * n + o.peer];
        push(o.lane, t);
        break;
    }
  }
}

// dst[0:bytes] = src, the work split into one 16-byte-aligned slice per lane
void UkComm::copy_sliced(char* dst, const char* src, uint64_t bytes) {
  const int L = cfg_.nlanes;
  const uint64_t per = round_up((bytes + L - 1) / L, 16);
  for (int l = 0; l < L; ++l) {
    const uint64_t lo = (uint64_t)l * per;
    if (lo >= bytes) break;
    UkTask t;
    memset(&t, 0, sizeof(t));
    t.op = UK_COPY;
    t.dst = (uint64_t)(dst + lo);
    t.src = (uint64_t)(src + lo);
    t.bytes = std::min(per, bytes - lo);
    push(l, t);
  }
}

uint64_t UkComm::all_reduce(const void* in, void* out, size_t count, int dtype, int op, UkAlgo algo,
                            cudaStream_t stream, bool symmetric) {
  UB_CHECK(dtype >= 0 && dtype < kNumDTypes, "ukernel all_reduce: bad dtype %d", dtype);
  UB_CHECK(op == kSum || op == kProd || op == kMax || op == kMin, "ukernel all_reduce: op %d unsupported", op);
  const uint64_t es = dtype_size(dtype), bytes = count * es;
  begin_op(stream);
  UkPlanParams p;
  p.nranks = comm_->nranks(), p.rank = comm_->rank(), p.nlanes = cfg_.nlanes;
  p.tile_bytes = cfg_.tile_bytes, p.elem_size = es, p.algo = algo;
  const bool zero_copy = symmetric && bytes > 0 && comm_->in_heap(in, bytes) && comm_->in_heap(out, bytes) &&
                         (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
  if (zero_copy) {
    ++stats_.zero_copy_ops;
    ++stats_.segments;
    run_plan(uk_plan_allreduce(bytes, p), Bufs{(char*)in, (char*)out}, dtype, op);
  } else {
    for (uint64_t off = 0; off < bytes; off += cfg_.staging_bytes) {
      const uint64_t seg = std::min<uint64_t>(cfg_.staging_bytes, bytes - off);
      ++stats_.segments;
      copy_sliced(stage_in_, (const char*)in + off, seg);
      lane_barrier();
      run_plan(uk_plan_allreduce(seg, p), Bufs{stage_in_, stage_out_}, dtype, op);
      lane_barrier();
      copy_sliced((char*)out + off, stage_out_, seg);
    }
  }
  return end_op(stream);
}

uint64_t UkComm::run_custom(const UkPlan& plan, const void* in, uint64_t in_bytes, void* out, uint64_t out_bytes,
                            int dtype, int op, cudaStream_t stream, bool symmetric) {
  UB_CHECK(dtype >= 0 && dtype < kNumDTypes, "ukernel run_custom: bad dtype %d", dtype);
  UB_CHECK(op == kSum || op == kProd || op == kMax || op == kMin, "ukernel run_custom: op %d unsupported", op);
  UB_CHECK(plan.nranks == comm_->nranks() && plan.rank == comm_->rank(), "ukernel run_custom: plan of rank %d/%d on rank %d/%d",
           plan.rank, plan.nranks, comm_->rank(), comm_->nranks());
  const std::string why = uk_check_bounds(plan, in_bytes, out_bytes, scratch_capacity(), cfg_.nlanes, dtype_size(dtype));
  UB_CHECK(why.empty(), "ukernel run_custom: %s", why.c_str());
  begin_op(stream);
  const bool zero_copy = symmetric && comm_->in_heap(in, std::max<uint64_t>(in_bytes, 1)) &&
                         comm_->in_heap(out, std::max<uint64_t>(out_bytes, 1)) && (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
  UkPlanParams bp;
  bp.nranks = comm_->nranks(), bp.rank = comm_->rank(), bp.nlanes = cfg_.nlanes, bp.tile_bytes = cfg_.tile_bytes;
  // A user program may Send into ANY buffer of a peer, so -- unlike the built-in plans, whose structure guarantees
  // it -- nothing says the peer is done staging / still reading that buffer: fence the program with a rank barrier
  // on both sides (all lanes join through lane_barrier).
  auto rank_barrier = [&]() {
    lane_barrier();
    run_plan(uk_plan_barrier(bp), Bufs{stage_in_, stage_out_}, kU8, kSum);
    lane_barrier();
  };
  ++stats_.segments;
  if (zero_copy) {
    ++stats_.zero_copy_ops;
    rank_barrier();
    run_plan(plan, Bufs{(char*)in, (char*)out}, dtype, op);
    rank_barrier();
  } else {
    UB_CHECK(in_bytes <= cfg_.staging_bytes && out_bytes <= cfg_.staging_bytes,
             "ukernel run_custom: %lu / %lu bytes do not fit the staging buffers (%lu); use symmetric tensors",
             (unsigned long)in_bytes, (unsigned long)out_bytes, (unsigned long)cfg_.staging_bytes);
    const bool in_place = (const void*)out == in;  // one staged buffer serves both names
    // the program may read Out before writing it (in-place algorithms): stage both in
    const uint64_t stage_in_bytes = in_place ? std::max(in_bytes, out_bytes) : in_bytes;
    if (stage_in_bytes) copy_sliced(stage_in_, (const char*)in, stage_in_bytes);
    if (out_bytes && !in_place) copy_sliced(stage_out_, (const char*)out, out_bytes);
    rank_barrier();
    run_plan(plan, Bufs{stage_in_, in_place ? stage_in_ : stage_out_}, dtype, op);
    rank_barrier();
    if (out_bytes) copy_sliced((char*)out, in_place ? stage_in_ : stage_out_, out_bytes);
  }
  return end_op(stream);
}

uint64_t UkComm::all_to_all(const void* in, void* out, size_t count_per_peer, int dtype, cudaStream_t stream,
                            bool symmetric) {
  UB_CHECK(dtype >= 0 && dtype < kNumDTypes, "ukernel all_to_all: bad dtype %d", dtype);
  UB_CHECK(in != out, "ukernel all_to_all: in-place operation is not supported");
  const int n = comm_->nranks();
  const uint64_t block = count_per_peer * dtype_size(dtype);
  begin_op(stream);
  UkPlanParams p;
  p.nranks = n, p.rank = comm_->rank(), p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  const bool zero_copy = symmetric && block > 0 && comm_->in_heap(in, block * n) && comm_->in_heap(out, block * n) &&
                         (((uintptr_t)in | (uintptr_t)out | block) & 15) == 0;
  if (zero_copy) {
    ++stats_.zero_copy_ops;
    ++stats_.segments;
    run_plan(uk_plan_alltoall(block, p), Bufs{(char*)in, (char*)out}, dtype, kSum);
  } else {
    // segment the per-peer block so that n blocks fit the staging buffers
    const uint64_t seg_max = cfg_.staging_bytes / n / 16 * 16;
    for (uint64_t off = 0; off < block; off += seg_max) {
      const uint64_t seg = std::min<uint64_t>(seg_max, block - off);
      const uint64_t pitch = round_up(seg, 16);  // staged blocks are 16-byte aligned whatever the element size
      ++stats_.segments;
      for (int q = 0; q < n; ++q) {
        UkTask t;
        memset(&t, 0, sizeof(t));
        t.op = UK_COPY;
        t.dst = (uint64_t)(stage_in_ + (uint64_t)q * pitch);
        t.src = (uint64_t)((const char*)in + (uint64_t)q * block + off);
        t.bytes = seg;
        push(q % cfg_.nlanes, t);
      }
      lane_barrier();
      run_plan(uk_plan_alltoall(pitch, p), Bufs{stage_in_, stage_out_}, dtype, kSum);
      lane_barrier();
      for (int q = 0; q < n; ++q) {
        UkTask t;
        memset(&t, 0, sizeof(t));
        t.op = UK_COPY;
        t.dst = (uint64_t)((char*)out + (uint64_t)q * block + off);
        t.src = (uint64_t)(stage_out_ + (uint64_t)q * pitch);
        t.bytes = seg;
        push(q % cfg_.nlanes, t);
      }
    }
  }
  return end_op(stream);
}

uint64_t UkComm::all_gather(const void* in, void* out, size_t count_per_rank, int dtype, cudaStream_t stream,
                            bool symmetric) {
  UB_CHECK(dtype >= 0 && dtype < kNumDTypes, "ukernel all_gather: bad dtype %d", dtype);
  const int n = comm_->nranks();
  const uint64_t block = count_per_rank * dtype_size(dtype);
  begin_op(stream);
  UkPlanParams p;
  p.nranks = n, p.rank = comm_->rank(), p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  const bool zero_copy = symmetric && block > 0 && comm_->in_heap(in, block) && comm_->in_heap(out, block * n) &&
                         (((uintptr_t)in | (uintptr_t)out | block) & 15) == 0;
  if (zero_copy) {
    ++stats_.zero_copy_ops;
    ++stats_.segments;
    run_plan(uk_plan_allgather(block, p), Bufs{(char*)in, (char*)out}, dtype, kSum);
  } else {
    const uint64_t seg_max = cfg_.staging_bytes / n / 16 * 16;
    for (uint64_t off = 0; off < block; off += seg_max) {
      const uint64_t seg = std::min<uint64_t>(seg_max, block - off);
      const uint64_t pitch = round_up(seg, 16);
      ++stats_.segments;
      copy_sliced(stage_in_, (const char*)in + off, seg);
      lane_barrier();
      run_plan(uk_plan_allgather(pitch, p), Bufs{stage_in_, stage_out_}, dtype, kSum);
      lane_barrier();
      for (int q = 0; q < n; ++q) {
        UkTask t;
        memset(&t, 0, sizeof(t));
        t.op = UK_COPY;
        t.dst = (uint64_t)((char*)out + (uint64_t)q * block + off);
        t.src = (uint64_t)(stage_out_ + (uint64_t)q * pitch);
        t.bytes = seg;
        push(q % cfg_.nlanes, t);
      }
    }
  }
  return end_op(stream);
}

uint64_t UkComm::reduce_scatter(const void* in, void* out, size_t recv_count, int dtype, int op, cudaStream_t stream) {
  UB_CHECK(dtype >= 0 && dtype < kNumDTypes, "ukernel reduce_scatter: bad dtype %d", dtype);
  UB_CHECK(op == kSum || op == kProd || op == kMax || op == kMin, "ukernel reduce_scatter: op %d unsupported", op);
  const int n = comm_->nranks();
  const uint64_t es = dtype_size(dtype), block = recv_count * es;
  begin_op(stream);
  UkPlanParams p;
  p.nranks = n, p.rank = comm_->rank(), p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes, p.elem_size = es;
  // n pieces must fit the input staging buffer; pieces stay multiples of 16 bytes (and of the element size)
  const uint64_t seg_max = cfg_.staging_bytes / n / 16 * 16;
  for (uint64_t off = 0; off < block; off += seg_max) {
    const uint64_t seg = std::min<uint64_t>(seg_max, block - off);
    const uint64_t pitch = round_up(seg, 16);
    ++stats_.segments;
    for (int q = 0; q < n; ++q) {
      UkTask t;
      memset(&t, 0, sizeof(t));
      t.op = UK_COPY;
      t.dst = (uint64_t)(stage_in_ + (uint64_t)q * pitch);
      t.src = (uint64_t)((const char*)in + (uint64_t)q * block + off);
      t.bytes = seg;
      push(q % cfg_.nlanes, t);
    }
    lane_barrier();
    // the plan reduces `pitch` bytes per piece; the (< 16 byte) padding holds stale data that is never copied out
    run_plan(uk_plan_reduce_scatter(pitch, p), Bufs{stage_in_, stage_out_}, dtype, op);
    lane_barrier();
    copy_sliced((char*)out + off, stage_out_, seg);
  }
  return end_op(stream);
}

uint64_t UkComm::broadcast(const void* in, void* out, size_t count, int dtype, int root, cudaStream_t stream) {
  UB_CHECK(dtype >= 0 && dtype < kNumDTypes, "ukernel broadcast: bad dtype %d", dtype);
  UB_CHECK(root >= 0 && root < comm_->nranks(), "ukernel broadcast: bad root %d", root);
  const uint64_t bytes = count * dtype_size(dtype);
  begin_op(stream);
  UkPlanParams p;
  p.nranks = comm_->nranks(), p.rank = comm_->rank(), p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  for (uint64_t off = 0; off < bytes; off += cfg_.staging_bytes) {
    const uint64_t seg = std::min<uint64_t>(cfg_.staging_bytes, bytes - off);
    ++stats_.segments;
    if (comm_->rank() == root) copy_sliced(stage_in_, (const char*)in + off, seg);
    lane_barrier();
    run_plan(uk_plan_broadcast(seg, root, p), Bufs{stage_in_, stage_out_}, dtype, kSum);
    lane_barrier();
    copy_sliced((char*)out + off, stage_out_, seg);
  }
  return end_op(stream);
}

uint64_t UkComm::barrier(cudaStream_t stream) {
  begin_op(stream);
  UkPlanParams p;
  p.nranks = comm_->nranks(), p.rank = comm_->rank(), p.nlanes = cfg_.nlanes, p.tile_bytes = cfg_.tile_bytes;
  run_plan(uk_plan_barrier(p), Bufs{stage_in_, stage_out_}, kU8, kSum);
  return end_op(stream);
}

}  // namespace ub
