// A stand-in for libnccl used by nccl_fallback_test.cc: implements the entry points the drop-in forwards to with
// recognisable side effects (marker bytes in receive buffers, call counters) so that the forwarding logic can be
// tested without GPUs or a real NCCL.
#include <nccl.h>

#include <cstring>

static int g_counts[8];  // 0 allreduce 1 reduce 2 broadcast 3 reducescatter 4 allgather 5 send 6 recv 7 groups
struct FakeComm {
  int rank, nranks;
};

extern "C" {
#define EXP __attribute__((visibility("default")))
EXP ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  memset(id, 0, sizeof(*id));
  memcpy(id->internal, "FAKE-NCCL-ID", 12);
  return ncclSuccess;
}
EXP ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
  if (memcmp(id.internal, "FAKE-NCCL-ID", 12) != 0) return ncclInvalidArgument;  // rank 0's id must reach everybody
  *comm = (ncclComm_t) new FakeComm{rank, nranks};
  return ncclSuccess;
}
EXP ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  delete (FakeComm*)comm;
  return ncclSuccess;
}
EXP const char* ncclGetErrorString(ncclResult_t) { return "fake nccl error"; }
static size_t esz(ncclDataType_t dt) { return dt == ncclFloat || dt == ncclInt32 || dt == ncclUint32 ? 4 : (dt == ncclFloat64 || dt == ncclInt64 || dt == ncclUint64 ? 8 : (dt == ncclHalf || dt == ncclBfloat16 ? 2 : 1)); }
EXP ncclResult_t ncclAllReduce(const void*, void* r, size_t n, ncclDataType_t dt, ncclRedOp_t, ncclComm_t, cudaStream_t) {
  ++g_counts[0];
  memset(r, 0xA5, n * esz(dt));
  return ncclSuccess;
}
EXP ncclResult_t ncclReduce(const void*, void* r, size_t n, ncclDataType_t dt, ncclRedOp_t, int root, ncclComm_t c, cudaStream_t) {
  ++g_counts[1];
  if (((FakeComm*)c)->rank == root) memset(r, 0xA6, n * esz(dt));
  return ncclSuccess;
}
EXP ncclResult_t ncclBroadcast(const void*, void* r, size_t n, ncclDataType_t dt, int root, ncclComm_t, cudaStream_t) {
  ++g_counts[2];
  memset(r, 0xB0 + root, n * esz(dt));
  return ncclSuccess;
}
EXP ncclResult_t ncclReduceScatter(const void*, void* r, size_t n, ncclDataType_t dt, ncclRedOp_t, ncclComm_t, cudaStream_t) {
  ++g_counts[3];
  memset(r, 0xA7, n * esz(dt));
  return ncclSuccess;
}
EXP ncclResult_t ncclAllGather(const void*, void* r, size_t n, ncclDataType_t dt, ncclComm_t c, cudaStream_t) {
  ++g_counts[4];
  memset(r, 0xA8, n * esz(dt) * (size_t)((FakeComm*)c)->nranks);
  return ncclSuccess;
}
EXP ncclResult_t ncclSend(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) {
  ++g_counts[5];
  return ncclSuccess;
}
EXP ncclResult_t ncclRecv(void* r, size_t n, ncclDataType_t dt, int peer, ncclComm_t, cudaStream_t) {
  ++g_counts[6];
  memset(r, 0x50 + peer, n * esz(dt));
  return ncclSuccess;
}
EXP ncclResult_t ncclGroupStart() {
  ++g_counts[7];
  return ncclSuccess;
}
EXP ncclResult_t ncclGroupEnd() { return ncclSuccess; }
EXP void fake_nccl_counts(int* out) { memcpy(out, g_counts, sizeof(g_counts)); }
}
