// C-style API of the transfer engine for external plugins (NIXL backend style).
// Function set mirrors the reference's p2p/uccl_engine.h:33-285 (create/connect/accept/reg/
// read/write/send/recv(+vector)/xfer_status/metadata/notifs/prepare_fifo/update_fifo/
// conn_is_local/get_ipc_info/update_ipc_info); the descriptor ("fifo item" / "ipc info") is this
// library's 128-byte XferDesc.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

typedef struct uccl_engine uccl_engine_t;
typedef struct uccl_conn uccl_conn_t;
typedef uint64_t uccl_mr_t;

#define UCCL_ENGINE_DESC_BYTES 128

typedef struct notify_msg {
  char name[64];
  char msg[1024];
} notify_msg_t;

uccl_engine_t* uccl_engine_create(int num_cpus, bool in_python);
uccl_engine_t* uccl_engine_create_on(int gpu_idx, int num_cpus);
void uccl_engine_destroy(uccl_engine_t* engine);
uccl_conn_t* uccl_engine_connect(uccl_engine_t* engine, char const* ip_addr, int remote_gpu_idx, int remote_port);
uccl_conn_t* uccl_engine_accept(uccl_engine_t* engine, char* ip_addr_buf, size_t ip_addr_buf_len, int* remote_gpu_idx);
int uccl_engine_start_listener(uccl_conn_t* conn);
void uccl_engine_stop_accept(uccl_engine_t* engine);
void uccl_engine_conn_destroy(uccl_conn_t* conn);
int uccl_engine_reg(uccl_engine_t* engine, uintptr_t data, size_t size, uccl_mr_t& mr_id);
void uccl_engine_mr_destroy(uccl_engine_t* engine, uccl_mr_t mr);
// one-sided ops take the remote window descriptor produced by prepare_fifo / get_ipc_info on the peer
int uccl_engine_read(uccl_conn_t* conn, uccl_mr_t mr, void const* data, size_t size, void const* remote_desc,
                     uint64_t* transfer_id);
int uccl_engine_write(uccl_conn_t* conn, uccl_mr_t mr, void const* data, size_t size, void const* remote_desc,
                      uint64_t* transfer_id);
int uccl_engine_read_vector(uccl_conn_t* conn, std::vector<uccl_mr_t> mr_ids, std::vector<void*> dst_v,
                            std::vector<size_t> size_v, std::vector<std::string> remote_descs, int num_iovs,
                            uint64_t* transfer_id);
int uccl_engine_write_vector(uccl_conn_t* conn, std::vector<uccl_mr_t> mr_ids, std::vector<void*> src_v,
                             std::vector<size_t> size_v, std::vector<std::string> remote_descs, int num_iovs,
                             uint64_t* transfer_id);
int uccl_engine_send(uccl_conn_t* conn, uccl_mr_t mr, void const* data, size_t size, uint64_t* transfer_id);
int uccl_engine_send_vector(uccl_conn_t* conn, std::vector<uccl_mr_t> mr_ids, std::vector<void const*> src_v,
                            std::vector<size_t> size_v, int num_iovs, uint64_t* transfer_id);
int uccl_engine_recv(uccl_conn_t* conn, uccl_mr_t mr, void* data, size_t size);
bool uccl_engine_xfer_status(uccl_conn_t* conn, uint64_t transfer_id);
int uccl_engine_get_metadata(uccl_engine_t* engine, char** metadata_str);
std::vector<notify_msg_t> uccl_engine_get_notifs();
int uccl_engine_send_notif(uccl_conn_t* conn, notify_msg_t* notify_msg);
int uccl_engine_prepare_fifo(uccl_engine_t* engine, uccl_mr_t mr, void const* data, size_t size, char* fifo_buf);
int uccl_engine_update_fifo(char* fifo_buf, uint64_t remote_addr, uint32_t size);
bool uccl_engine_conn_is_local(uccl_conn_t* conn);
int uccl_engine_get_ipc_info(uccl_engine_t* engine, uintptr_t addr, char* ipc_buf, bool* has_ipc);
int uccl_engine_update_ipc_info(char* ipc_buf, uintptr_t addr, size_t size);

// ---------------------------------------------------------------------------------------------------------------
// Source compatibility with plugins written against the reference's header (p2p/uccl_engine.h + include/common.h):
// the same macro and type names, and overloads with the reference's parameter lists, as inline wrappers over the
// functions above.  The window descriptor ("fifo item") of this library is 128 bytes (it carries a CUDA IPC handle),
// so FIFO_SIZE is 128 here and FifoItem is an opaque block of that size: code that sizes its buffers with the macros
// and fills items through prepare_fifo / update_fifo / deserialize_fifo_item compiles and works unchanged; code that
// reads FifoItem fields (addr / rkey) is RDMA-specific and has no meaning on NVLink.
#ifndef MSG_SIZE
#define MSG_SIZE 256
#endif
#ifndef FIFO_SIZE
#define FIFO_SIZE UCCL_ENGINE_DESC_BYTES
#endif
#ifndef IPC_INFO_SIZE
#define IPC_INFO_SIZE UCCL_ENGINE_DESC_BYTES
#endif

struct FifoItem {
  char raw[UCCL_ENGINE_DESC_BYTES];
};
inline void serialize_fifo_item(FifoItem const& item, char* buf) { __builtin_memcpy(buf, item.raw, sizeof(item.raw)); }
inline void deserialize_fifo_item(char const* buf, FifoItem* item) { __builtin_memcpy(item->raw, buf, sizeof(item->raw)); }

// remote_gpu: a CUDA device index as a string ("0") like the reference accepts; a PCI BDF string cannot be mapped to
// an index of another host here and is passed on as "unknown" (-1) -- the index is informational for this engine
inline uccl_conn_t* uccl_engine_connect(uccl_engine_t* engine, char const* ip_addr, char const* remote_gpu, int remote_port,
                                        bool same_process = false) {
  (void)same_process;
  int idx = -1;
  if (remote_gpu && *remote_gpu) {
    bool digits = true;
    for (char const* p = remote_gpu; *p; ++p) digits = digits && *p >= '0' && *p <= '9';
    if (digits) idx = atoi(remote_gpu);
  }
  return uccl_engine_connect(engine, ip_addr, idx, remote_port);
}
inline int uccl_engine_read(uccl_conn_t* conn, uccl_mr_t mr, void const* data, size_t size, FifoItem fifo_item,
                            uint64_t* transfer_id) {
  return uccl_engine_read(conn, mr, data, size, (void const*)fifo_item.raw, transfer_id);
}
inline int uccl_engine_write(uccl_conn_t* conn, uccl_mr_t mr, void const* data, size_t size, FifoItem fifo_item,
                             uint64_t* transfer_id) {
  return uccl_engine_write(conn, mr, data, size, (void const*)fifo_item.raw, transfer_id);
}
inline std::vector<std::string> uccl_engine_descs_of(std::vector<FifoItem> const& items, std::vector<char*> const& ipc_bufs) {
  std::vector<std::string> out;
  for (size_t i = 0; i < items.size(); ++i)  // a per-block IPC info buffer, when given, is the same 128-byte descriptor
    out.emplace_back(i < ipc_bufs.size() && ipc_bufs[i] ? ipc_bufs[i] : items[i].raw, (size_t)UCCL_ENGINE_DESC_BYTES);
  return out;
}
inline int uccl_engine_read_vector(uccl_conn_t* conn, std::vector<uccl_mr_t> mr_ids, std::vector<void*> dst_v,
                                   std::vector<size_t> size_v, std::vector<FifoItem> fifo_items, int num_iovs,
                                   uint64_t* transfer_id, std::vector<char*> ipc_bufs = {}) {
  return uccl_engine_read_vector(conn, mr_ids, dst_v, size_v, uccl_engine_descs_of(fifo_items, ipc_bufs), num_iovs, transfer_id);
}
inline int uccl_engine_write_vector(uccl_conn_t* conn, std::vector<uccl_mr_t> mr_ids, std::vector<void*> src_v,
                                    std::vector<size_t> size_v, std::vector<FifoItem> fifo_items, int num_iovs,
                                    uint64_t* transfer_id, std::vector<char*> ipc_bufs = {}) {
  return uccl_engine_write_vector(conn, mr_ids, src_v, size_v, uccl_engine_descs_of(fifo_items, ipc_bufs), num_iovs, transfer_id);
}
inline int uccl_engine_update_fifo(FifoItem& fifo_item, uint64_t remote_addr, uint32_t size) {
  return uccl_engine_update_fifo(fifo_item.raw, remote_addr, size);
}
// base_addr (the registered region's base) is implied by the descriptor itself here
inline int uccl_engine_update_ipc_info(char* ipc_buf, uintptr_t addr, uintptr_t base_addr, size_t size) {
  (void)base_addr;
  return uccl_engine_update_ipc_info(ipc_buf, addr, size);
}
