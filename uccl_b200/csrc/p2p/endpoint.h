// NIXL-style initiator/target transfer engine for one NVSwitch node.
//
// API parity target: the reference's `class Endpoint` (p2p/engine.h:246-552) -- connection,
// memory-registration and transfer id tables; blocking + async send/recv (+vector), one-sided
// read/write (+vector) against advertised descriptors, async handles polled to completion,
// metadata exchange, notifications.  What differs by design: there is exactly one transport
// (peer HBM mapped through CUDA IPC, moved by the in-kernel TMA copy engine of
// p2p_kernels.cu on side streams), so the RDMA/TCP/NCCL backends, shm jring mailboxes and
// per-op cudaIpcOpen/Close of the reference disappear; control messages ride one TCP
// connection per (unidirectional) conn.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "p2p_types.h"

namespace ub {

struct P2PStats {
  uint64_t bytes_sent = 0, bytes_received = 0, bytes_read = 0, bytes_written = 0;
  uint64_t transfers = 0, kernel_launches = 0, memcpy_fallbacks = 0;
};

class Endpoint {
 public:
  explicit Endpoint(int local_gpu_idx, int num_streams = 4);
  ~Endpoint();
  Endpoint(const Endpoint&) = delete;

  // ---- identity / metadata (ip, port, gpu index)
  std::string get_metadata() const;
  static bool parse_metadata(const std::string& md, std::string* ip, uint16_t* port, int* gpu_idx);
  int gpu_idx() const { return gpu_; }
  uint16_t port() const { return port_; }

  // ---- connections (a conn is initiator -> target; both sides may use it in both directions)
  bool connect(const std::string& ip, int remote_gpu_idx, uint16_t remote_port, uint64_t* conn_id);
  bool accept(std::string* ip, int* remote_gpu_idx, uint64_t* conn_id, int timeout_ms = -1);
  bool add_remote_endpoint(const std::string& metadata, uint64_t* conn_id);
  bool remove_remote_endpoint(uint64_t conn_id);
  bool start_passive_accept() { return true; }  // the accept thread always runs
  bool conn_is_local(uint64_t) const { return true; }

  // ---- memory registration
  bool reg(const void* ptr, size_t size, uint64_t* mr_id);
  bool dereg(uint64_t mr_id);
  bool describe(const void* ptr, size_t size, XferDesc* out);  // window descriptor for any registered/unregistered ptr

  // ---- two-sided
  bool send_async(uint64_t conn, const std::vector<const void*>& ptrs, const std::vector<size_t>& sizes, uint64_t* tid);
  bool recv_async(uint64_t conn, const std::vector<void*>& ptrs, const std::vector<size_t>& sizes, uint64_t* tid);
  // ---- one-sided against advertised descriptors
  bool write_async(uint64_t conn, const std::vector<const void*>& src, const std::vector<size_t>& sizes,
                   const std::vector<XferDesc>& remote, uint64_t* tid);
  bool read_async(uint64_t conn, const std::vector<void*>& dst, const std::vector<size_t>& sizes,
                  const std::vector<XferDesc>& remote, uint64_t* tid);
  bool advertise(uint64_t conn, const void* ptr, size_t size, XferDesc* out);
  // ---- prepared transfers (NIXL's prepXfer / postXfer): descriptors are resolved, peers mapped and the copy kernel's
  //      descriptor tables built ONCE; every post() is then a bare kernel launch + event, however many blocks it moves
  //      (a KV-cache mover re-sends the same page lists).  Load/store reachable peers only (returns false otherwise).
  bool prepare(uint64_t conn, bool is_write, const std::vector<const void*>& local, const std::vector<size_t>& sizes,
               const std::vector<XferDesc>& remote, uint64_t* prep_id);
  bool post(uint64_t prep_id, uint64_t* tid);
  bool release(uint64_t prep_id);

  // poll once: *done = finished (the handle is released when done, like the reference's poll_async)
  bool poll_async(uint64_t tid, bool* done);
  bool wait(uint64_t tid, int timeout_ms = -1);

  // ---- notifications (uccl_engine_send_notif / get_notifs)
  bool send_notif(uint64_t conn, const std::string& msg);
  std::vector<std::pair<uint64_t, std::string>> get_notifs();

  P2PStats stats() const;

 private:
  struct Conn;
  struct Transfer;
  void engine_loop();
  void handle_message(Conn& c, uint32_t type, uint64_t seq, std::vector<char>& payload);
  bool send_msg(Conn& c, uint32_t type, uint64_t seq, const void* payload, uint32_t len);
  bool send_msg2(Conn& c, uint32_t type, uint64_t seq, const void* p1, uint32_t len1, const void* p2, uint32_t len2);
  bool tcp_write_buffers(Conn& c, const std::vector<const char*>& src, const std::vector<uint64_t>& dst_addr,
                         const std::vector<size_t>& sizes);
  bool remote_is_other_process(const XferDesc& d) const;  // not load/store reachable: payload goes over the connection
  void copy_any(void* dst, const void* src, size_t n);    // memcpy in host mode, cudaMemcpy(Default) with a GPU
  void run_helper(std::function<void()> fn);
  void progress_locked();
  bool launch_copy(const std::vector<const char*>& src, const std::vector<char*>& dst,
                   const std::vector<size_t>& sizes, cudaEvent_t ev);
  struct DescTable;
  struct CopyLaunch {  // one launch of the copy kernel
    P2PCopyBatch b;
    int grid = 1;
    DescTable* tab = nullptr;
  };
  bool is_device_ptr(const void* p);  // cached cudaPointerGetAttributes (2 MiB granules)
  // tables are uploaded on `st` (stream-ordered before the launches that read them)
  bool build_launches(const std::vector<const char*>& src, const std::vector<char*>& dst, const std::vector<size_t>& sizes,
                      std::vector<CopyLaunch>* out, cudaStream_t st);
  struct Prepared {
    std::shared_ptr<Conn> conn;
    bool is_write = true;
    uint64_t bytes = 0;
    std::vector<CopyLaunch> launches;
    // a few very large blocks: the copy engines move them (same policy as launch_copy)
    bool use_memcpy = false;
    std::vector<const char*> src;
    std::vector<char*> dst;
    std::vector<size_t> sizes;
  };
  std::map<uint64_t, std::shared_ptr<Prepared>> prepared_;
  uint64_t next_prep_ = 1;
  std::unordered_map<uint64_t, bool> ptr_is_device_;
  void* map_remote(const XferDesc& d);
  std::shared_ptr<Conn> find_conn(uint64_t id);
  void wake();

  int gpu_;
  uint16_t port_ = 0;
  std::string ip_;
  int listen_fd_ = -1;
  int wake_fd_ = -1;
  std::atomic<bool> stop_{false};
  std::thread engine_;
  // host mode: the windows this process has handed to peers (registered / advertised / posted receives).
  // The TCP data path only touches memory inside one of them -- a peer cannot aim MSG_WRITE / READ_REQ at
  // an arbitrary address of this process.
  std::mutex exp_mu_;
  std::vector<std::pair<uint64_t, uint64_t>> exposed_;
  void expose(uint64_t addr, uint64_t size);
  bool is_exposed(uint64_t addr, uint64_t n);
  mutable std::mutex mu_;  // guards every table below
  std::condition_variable accept_cv_;
  std::map<uint64_t, std::shared_ptr<Conn>> conns_;
  std::vector<std::shared_ptr<Conn>> closing_;  // removed connections whose descriptor the engine thread still has to close
  void retire_closed_locked();
  std::deque<uint64_t> accepted_;
  std::map<uint64_t, std::shared_ptr<Transfer>> transfers_;
  std::deque<std::pair<uint64_t, std::string>> notifs_;
  uint64_t next_conn_ = 1, next_tid_ = 1, next_mr_ = 1;
  struct MR {
    const void* ptr;
    size_t size;
  };
  std::map<uint64_t, MR> mrs_;
  std::unordered_map<std::string, void*> ipc_open_;  // handle bytes -> mapped base
  std::unordered_map<uint64_t, std::string> ipc_export_;  // allocation base -> handle bytes
  std::vector<cudaStream_t> streams_;
  size_t next_stream_ = 0;
  std::vector<cudaEvent_t> event_pool_;
  // pinned, device-mapped descriptor tables for batches of more than kP2PMaxEntries blocks; a table goes back
  // to the pool from a stream callback once its kernel has finished
  struct DescTable {
    void* host = nullptr;  // pinned staging: [kP2PTableEntries] P2PCopyEntry, then [kP2PTableEntries + 1] u32 chunk prefix
    void* dev = nullptr;   // the device copy the kernel reads
    Endpoint* owner = nullptr;
  };
  std::vector<DescTable*> table_pool_;
  std::vector<DescTable*> tables_all_;
  DescTable* acquire_table();
  static void release_table_cb(void* p);
  P2PStats stats_;
  uint32_t peer_enabled_mask_ = 0;
  std::mutex helpers_mu_;
  std::vector<std::thread> helpers_;  // blocking TCP payload sends never run on the engine thread
};

}  // namespace ub
