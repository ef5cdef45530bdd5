#include "nccl_fallback.h"

#include <dlfcn.h>

#include <cstring>
#include <sstream>
#include <stdexcept>

#include "../common/log.h"
#include "../common/param.h"

namespace ub {

UB_PARAM(FallbackMinBytes, "NCCL_FALLBACK_MIN_BYTES", 0)

struct NcclFallback::Fns {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(void**, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(void*);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, void*, cudaStream_t);
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, void*, cudaStream_t);
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, void*, cudaStream_t);
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, void*, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, void*, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, void*, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, void*, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
};

unsigned NcclFallback::parse_ops(const std::string& list) {
  unsigned ops = 0;
  std::stringstream ss(list);
  std::string tok;
  while (std::getline(ss, tok, ',')) {
    std::string t;
    for (char ch : tok)
      if (ch != ' ' && ch != '_' && ch != '-') t.push_back((char)tolower((unsigned char)ch));
    if (t.empty() || t == "none") continue;
    if (t == "allreduce") ops |= kAllReduce;
    else if (t == "reduce") ops |= kReduce;
    else if (t == "broadcast" || t == "bcast") ops |= kBroadcast;
    else if (t == "reducescatter") ops |= kReduceScatter;
    else if (t == "allgather") ops |= kAllGather;
    else if (t == "sendrecv" || t == "send" || t == "recv" || t == "p2p") ops |= kSendRecv;
    else if (t == "all") ops |= kAllReduce | kReduce | kBroadcast | kReduceScatter | kAllGather | kSendRecv;
    else throw std::runtime_error("uccl_b200: unknown operation '" + tok + "' in UCCL_B200_NCCL_FALLBACK_OPS");
  }
  return ops;
}

std::unique_ptr<NcclFallback> NcclFallback::create(int rank, int nranks, const std::function<void(void*)>& share_id) {
  const std::string path = param_load_str("NCCL_FALLBACK_LIB", "");
  const unsigned ops = parse_ops(param_load_str("NCCL_FALLBACK_OPS", ""));
  if (path.empty() || ops == 0) return nullptr;
  std::unique_ptr<NcclFallback> fb(new NcclFallback());
  fb->ops_ = ops;
  fb->min_bytes_ = (size_t)std::max<int64_t>(0, ubParamFallbackMinBytes());
  fb->lib_ = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
  if (!fb->lib_) throw std::runtime_error("uccl_b200: cannot load UCCL_B200_NCCL_FALLBACK_LIB=" + path + ": " + dlerror());
  fb->f_ = new Fns();
  auto need = [&](const char* name) -> void* {
    void* p = dlsym(fb->lib_, name);
    if (!p) throw std::runtime_error(std::string("uccl_b200: ") + path + " has no symbol " + name);
    return p;
  };
#define UB_FN(field, sym) fb->f_->field = reinterpret_cast<decltype(fb->f_->field)>(need(sym))
  UB_FN(GetUniqueId, "ncclGetUniqueId");
  UB_FN(CommInitRank, "ncclCommInitRank");
  UB_FN(CommDestroy, "ncclCommDestroy");
  UB_FN(GetErrorString, "ncclGetErrorString");
  UB_FN(AllReduce, "ncclAllReduce");
  UB_FN(Reduce, "ncclReduce");
  UB_FN(Broadcast, "ncclBroadcast");
  UB_FN(ReduceScatter, "ncclReduceScatter");
  UB_FN(AllGather, "ncclAllGather");
  UB_FN(Send, "ncclSend");
  UB_FN(Recv, "ncclRecv");
  UB_FN(GroupStart, "ncclGroupStart");
  UB_FN(GroupEnd, "ncclGroupEnd");
#undef UB_FN
  ncclUniqueId id;
  memset(&id, 0, sizeof(id));
  if (rank == 0) {
    ncclResult_t r = fb->f_->GetUniqueId(&id);
    if (r != ncclSuccess) throw std::runtime_error(std::string("uccl_b200: fallback ncclGetUniqueId: ") + fb->f_->GetErrorString(r));
  }
  share_id(&id);  // collective: every rank now holds rank 0's id
  ncclResult_t r = fb->f_->CommInitRank(&fb->comm_, nranks, id, rank);
  if (r != ncclSuccess)
    throw std::runtime_error(std::string("uccl_b200: fallback ncclCommInitRank: ") + fb->f_->GetErrorString(r));
  UB_INFO(SUB_INIT, "NCCL fallback: %s for ops mask 0x%x, collectives >= %zu bytes (rank %d/%d)", path.c_str(), ops,
          fb->min_bytes_, rank, nranks);
  return fb;
}

NcclFallback::~NcclFallback() {
  if (comm_ && f_) f_->CommDestroy(comm_);
  delete f_;
  if (lib_) dlclose(lib_);
}

ncclResult_t NcclFallback::all_reduce(const void* s, void* r, size_t n, ncclDataType_t dt, ncclRedOp_t op, cudaStream_t st) {
  ++forwarded_;
  return f_->AllReduce(s, r, n, dt, op, comm_, st);
}
ncclResult_t NcclFallback::reduce(const void* s, void* r, size_t n, ncclDataType_t dt, ncclRedOp_t op, int root,
                                  cudaStream_t st) {
  ++forwarded_;
  return f_->Reduce(s, r, n, dt, op, root, comm_, st);
}
ncclResult_t NcclFallback::broadcast(const void* s, void* r, size_t n, ncclDataType_t dt, int root, cudaStream_t st) {
  ++forwarded_;
  return f_->Broadcast(s, r, n, dt, root, comm_, st);
}
ncclResult_t NcclFallback::reduce_scatter(const void* s, void* r, size_t n, ncclDataType_t dt, ncclRedOp_t op,
                                          cudaStream_t st) {
  ++forwarded_;
  return f_->ReduceScatter(s, r, n, dt, op, comm_, st);
}
ncclResult_t NcclFallback::all_gather(const void* s, void* r, size_t n, ncclDataType_t dt, cudaStream_t st) {
  ++forwarded_;
  return f_->AllGather(s, r, n, dt, comm_, st);
}
ncclResult_t NcclFallback::send(const void* s, size_t n, ncclDataType_t dt, int peer, cudaStream_t st) {
  ++forwarded_;
  return f_->Send(s, n, dt, peer, comm_, st);
}
ncclResult_t NcclFallback::recv(void* r, size_t n, ncclDataType_t dt, int peer, cudaStream_t st) {
  ++forwarded_;
  return f_->Recv(r, n, dt, peer, comm_, st);
}
ncclResult_t NcclFallback::group_start() { return f_->GroupStart(); }
ncclResult_t NcclFallback::group_end() { return f_->GroupEnd(); }

}  // namespace ub
