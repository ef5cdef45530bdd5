// C-style API of the transfer engine for external plugins (NIXL backend style).
// Function set mirrors the reference's p2p/uccl_engine.h:33-285 (create/connect/accept/reg/
// read/write/send/recv(+vector)/xfer_status/metadata/notifs/prepare_fifo/update_fifo/
// conn_is_local/get_ipc_info/update_ipc_info); the descriptor ("fifo item" / "ipc info") is this
// library's 128-byte XferDesc.
#pragma once
#include <stddef.h>
#include <stdint.h>

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
