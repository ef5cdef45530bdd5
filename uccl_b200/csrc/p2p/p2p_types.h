// Host-safe structs of the P2P copy kernel and wire descriptors.
#pragma once
#include <stdint.h>

namespace ub {

constexpr int kP2PStages = 4;
constexpr int kP2PMaxEntries = 64;

struct P2PCopyEntry {
  const char* src;
  char* dst;
  uint64_t bytes;
  uint64_t bulk_bytes;  // 16-byte aligned prefix moved by TMA (0 if src/dst are not 16-byte aligned)
};

struct P2PCopyBatch {
  P2PCopyEntry e[kP2PMaxEntries];
  uint32_t chunk_prefix[kP2PMaxEntries + 1];  // exclusive prefix of per-entry bulk chunk counts
  int n;
  uint32_t chunk_bytes;
};

// A registered/advertised memory window, shipped between endpoints (128 bytes on the wire).
// Role of the reference's FifoItem / IpcTransferInfo (p2p/include/common.h:40-49, p2p/engine.h:82-90).
struct XferDesc {
  unsigned char ipc_handle[64];  // cudaIpcMemHandle_t of the allocation base
  uint64_t base;                 // allocation base VA in the owner process
  uint64_t addr;                 // window start VA in the owner process
  uint64_t size;                 // window bytes
  int32_t pid;                   // owner process
  int32_t dev;                   // owner CUDA device
  uint32_t kind;                 // 0 = CUDA IPC, 1 = pinned host (same process only), 2 = same-process device ptr
  uint32_t mr_id;
  uint64_t host_id;              // which machine owns the window (0 = unknown / legacy): other hosts use the wire
  unsigned char pad[16];
};
static_assert(sizeof(XferDesc) == 128, "XferDesc must be 128 bytes");

}  // namespace ub
