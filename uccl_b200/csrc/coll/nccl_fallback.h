// Optional forwarding of selected NCCL entry points to a real libnccl (dlopen), the escape hatch the reference's shim
// is built around (experimental/lite/nccl/nccl.cu:707-860: `MSCCLPP_NCCL_LIB_PATH`, per-operation force list
// `MSCCLPP_FORCE_NCCL_FALLBACK_OPERATION` :1866-1872).  Every operation is native here, so nothing is forwarded by
// default; a deployment that wants NCCL for an operation -- e.g. send/recv or broadcast of >= 256 MiB ordinary
// buffers, where the staged kernels are behind (ROADMAP) -- names it:
//
//   UCCL_B200_NCCL_FALLBACK_LIB=/usr/lib/x86_64-linux-gnu/libnccl.so.2
//   UCCL_B200_NCCL_FALLBACK_OPS=sendrecv,broadcast      (allreduce, reduce, broadcast, reducescatter, allgather, sendrecv)
//   UCCL_B200_NCCL_FALLBACK_MIN_BYTES=268435456         (collectives below this size stay native; send/recv ignores it)
//
// The library is loaded with RTLD_LOCAL | RTLD_DEEPBIND so that its own internal references do not resolve to the
// nccl* symbols this drop-in exports.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <functional>
#include <memory>
#include <string>

namespace ub {

class NcclFallback {
 public:
  enum Op : unsigned { kAllReduce = 1, kReduce = 2, kBroadcast = 4, kReduceScatter = 8, kAllGather = 16, kSendRecv = 32 };
  // `share_id(buf)`: rank 0 passes its 128-byte id in, every rank leaves with rank 0's (collective).
  // Returns nullptr when no fallback is configured; throws std::runtime_error when it is configured but unusable.
  static std::unique_ptr<NcclFallback> create(int rank, int nranks, const std::function<void(void*)>& share_id);
  ~NcclFallback();
  bool takes(Op op, size_t bytes) const { return (ops_ & op) && (op == kSendRecv || bytes >= min_bytes_); }
  static unsigned parse_ops(const std::string& list);  // throws on an unknown name

  ncclResult_t all_reduce(const void* s, void* r, size_t n, ncclDataType_t dt, ncclRedOp_t op, cudaStream_t st);
  ncclResult_t reduce(const void* s, void* r, size_t n, ncclDataType_t dt, ncclRedOp_t op, int root, cudaStream_t st);
  ncclResult_t broadcast(const void* s, void* r, size_t n, ncclDataType_t dt, int root, cudaStream_t st);
  ncclResult_t reduce_scatter(const void* s, void* r, size_t n, ncclDataType_t dt, ncclRedOp_t op, cudaStream_t st);
  ncclResult_t all_gather(const void* s, void* r, size_t n, ncclDataType_t dt, cudaStream_t st);
  ncclResult_t send(const void* s, size_t n, ncclDataType_t dt, int peer, cudaStream_t st);
  ncclResult_t recv(void* r, size_t n, ncclDataType_t dt, int peer, cudaStream_t st);
  ncclResult_t group_start();
  ncclResult_t group_end();
  uint64_t forwarded() const { return forwarded_; }

 private:
  NcclFallback() = default;
  void* lib_ = nullptr;
  void* comm_ = nullptr;  // the real library's ncclComm_t
  unsigned ops_ = 0;
  size_t min_bytes_ = 0;
  uint64_t forwarded_ = 0;
  struct Fns;
  Fns* f_ = nullptr;
};

}  // namespace ub
