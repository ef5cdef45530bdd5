#include "uccl_engine.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>

#include "../common/log.h"
#include "endpoint.h"

using namespace ub;

struct uccl_engine {
  std::shared_ptr<Endpoint> ep;
};
struct uccl_conn {
  uccl_engine* engine;
  uint64_t id;
};

namespace {
std::mutex g_mu;
std::vector<uccl_engine*> g_engines;
std::map<uint64_t, std::pair<const void*, size_t>> g_mrs;
#define UB_VISIBLE __attribute__((visibility("default")))
}  // namespace

UB_VISIBLE uccl_engine_t* uccl_engine_create_on(int gpu_idx, int num_cpus) {
  try {
    auto* e = new uccl_engine();
    e->ep = std::make_shared<Endpoint>(gpu_idx, num_cpus > 0 ? num_cpus : 4);
    std::lock_guard<std::mutex> g(g_mu);
    g_engines.push_back(e);
    return e;
  } catch (const std::exception& ex) {
    UB_ERROR("uccl_engine_create failed: %s", ex.what());
    return nullptr;
  }
}

UB_VISIBLE uccl_engine_t* uccl_engine_create(int num_cpus, bool /*in_python*/) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return uccl_engine_create_on(dev, num_cpus);
}

UB_VISIBLE void uccl_engine_destroy(uccl_engine_t* engine) {
  if (!engine) return;
  {
    std::lock_guard<std::mutex> g(g_mu);
    for (auto it = g_engines.begin(); it != g_engines.end(); ++it)
      if (*it == engine) {
        g_engines.erase(it);
        break;
      }
  }
  delete engine;
}

UB_VISIBLE uccl_conn_t* uccl_engine_connect(uccl_engine_t* engine, char const* ip_addr, int remote_gpu_idx,
                                            int remote_port) {
  uint64_t id = 0;
  if (!engine || !engine->ep->connect(ip_addr, remote_gpu_idx, (uint16_t)remote_port, &id)) return nullptr;
  return new uccl_conn{engine, id};
}

UB_VISIBLE uccl_conn_t* uccl_engine_accept(uccl_engine_t* engine, char* ip_addr_buf, size_t ip_addr_buf_len,
                                           int* remote_gpu_idx) {
  std::string ip;
  int gpu = -1;
  uint64_t id = 0;
  if (!engine || !engine->ep->accept(&ip, &gpu, &id, -1)) return nullptr;
  if (ip_addr_buf && ip_addr_buf_len) snprintf(ip_addr_buf, ip_addr_buf_len, "%s", ip.c_str());
  if (remote_gpu_idx) *remote_gpu_idx = gpu;
  return new uccl_conn{engine, id};
}

UB_VISIBLE int uccl_engine_start_listener(uccl_conn_t*) { return 0; }  // the engine thread always listens
UB_VISIBLE void uccl_engine_stop_accept(uccl_engine_t*) {}

UB_VISIBLE void uccl_engine_conn_destroy(uccl_conn_t* conn) {
  if (!conn) return;
  conn->engine->ep->remove_remote_endpoint(conn->id);
  delete conn;
}

UB_VISIBLE int uccl_engine_reg(uccl_engine_t* engine, uintptr_t data, size_t size, uccl_mr_t& mr_id) {
  uint64_t id = 0;
  if (!engine || !engine->ep->reg((const void*)data, size, &id)) return -1;
  mr_id = id;
  std::lock_guard<std::mutex> g(g_mu);
  g_mrs[id] = {(const void*)data, size};
  return 0;
}

UB_VISIBLE void uccl_engine_mr_destroy(uccl_engine_t* engine, uccl_mr_t mr) {
  if (engine) engine->ep->dereg(mr);
  std::lock_guard<std::mutex> g(g_mu);
  g_mrs.erase(mr);
}

static int one_sided(uccl_conn_t* conn, bool write, std::vector<void*> local, std::vector<size_t> sizes,
                     std::vector<std::string> descs, uint64_t* tid) {
  if (!conn) return -1;
  std::vector<XferDesc> r(descs.size());
  for (size_t i = 0; i < descs.size(); ++i) {
    if (descs[i].size() != sizeof(XferDesc)) return -1;
    memcpy(&r[i], descs[i].data(), sizeof(XferDesc));
  }
  uint64_t t = 0;
  bool ok;
  if (write) {
    std::vector<const void*> src(local.begin(), local.end());
    ok = conn->engine->ep->write_async(conn->id, src, sizes, r, &t);
  } else {
    ok = conn->engine->ep->read_async(conn->id, local, sizes, r, &t);
  }
  if (!ok) return -1;
  if (tid) *tid = t;
  return 0;
}

UB_VISIBLE int uccl_engine_read(uccl_conn_t* conn, uccl_mr_t, void const* data, size_t size, void const* remote_desc,
                                uint64_t* transfer_id) {
  return one_sided(conn, false, {const_cast<void*>(data)}, {size},
                   {std::string((const char*)remote_desc, sizeof(XferDesc))}, transfer_id);
}
UB_VISIBLE int uccl_engine_write(uccl_conn_t* conn, uccl_mr_t, void const* data, size_t size, void const* remote_desc,
                                 uint64_t* transfer_id) {
  return one_sided(conn, true, {const_cast<void*>(data)}, {size},
                   {std::string((const char*)remote_desc, sizeof(XferDesc))}, transfer_id);
}
UB_VISIBLE int uccl_engine_read_vector(uccl_conn_t* conn, std::vector<uccl_mr_t>, std::vector<void*> dst_v,
                                       std::vector<size_t> size_v, std::vector<std::string> remote_descs, int,
                                       uint64_t* transfer_id) {
  return one_sided(conn, false, dst_v, size_v, remote_descs, transfer_id);
}
UB_VISIBLE int uccl_engine_write_vector(uccl_conn_t* conn, std::vector<uccl_mr_t>, std::vector<void*> src_v,
                                        std::vector<size_t> size_v, std::vector<std::string> remote_descs, int,
                                        uint64_t* transfer_id) {
  return one_sided(conn, true, src_v, size_v, remote_descs, transfer_id);
}

UB_VISIBLE int uccl_engine_send(uccl_conn_t* conn, uccl_mr_t, void const* data, size_t size, uint64_t* transfer_id) {
  if (!conn) return -1;
  uint64_t t = 0;
  if (!conn->engine->ep->send_async(conn->id, {data}, {size}, &t)) return -1;
  if (transfer_id) *transfer_id = t;
  return 0;
}
UB_VISIBLE int uccl_engine_send_vector(uccl_conn_t* conn, std::vector<uccl_mr_t>, std::vector<void const*> src_v,
                                       std::vector<size_t> size_v, int, uint64_t* transfer_id) {
  if (!conn) return -1;
  uint64_t t = 0;
  if (!conn->engine->ep->send_async(conn->id, src_v, size_v, &t)) return -1;
  if (transfer_id) *transfer_id = t;
  return 0;
}
UB_VISIBLE int uccl_engine_recv(uccl_conn_t* conn, uccl_mr_t, void* data, size_t size) {
  if (!conn) return -1;
  uint64_t t = 0;
  if (!conn->engine->ep->recv_async(conn->id, {data}, {size}, &t)) return -1;
  return conn->engine->ep->wait(t, -1) ? 0 : -1;
}
UB_VISIBLE bool uccl_engine_xfer_status(uccl_conn_t* conn, uint64_t transfer_id) {
  if (!conn) return false;
  bool done = false;
  if (!conn->engine->ep->poll_async(transfer_id, &done)) return true;  // already reaped
  return done;
}

UB_VISIBLE int uccl_engine_get_metadata(uccl_engine_t* engine, char** metadata_str) {
  if (!engine || !metadata_str) return -1;
  std::string ip;
  uint16_t port;
  int gpu;
  Endpoint::parse_metadata(engine->ep->get_metadata(), &ip, &port, &gpu);
  char buf[128];
  snprintf(buf, sizeof(buf), "%s:%u?%d", ip.c_str(), (unsigned)port, gpu);  // "ip:port?gpu" like the reference
  *metadata_str = strdup(buf);
  return 0;
}

UB_VISIBLE std::vector<notify_msg_t> uccl_engine_get_notifs() {
  std::vector<notify_msg_t> out;
  std::lock_guard<std::mutex> g(g_mu);
  for (auto* e : g_engines)
    for (auto& kv : e->ep->get_notifs()) {
      notify_msg_t m;
      memset(&m, 0, sizeof(m));
      const std::string& s = kv.second;
      size_t sep = s.find('\0');
      if (sep == std::string::npos) {
        snprintf(m.msg, sizeof(m.msg), "%s", s.c_str());
      } else {
        snprintf(m.name, sizeof(m.name), "%s", s.substr(0, sep).c_str());
        size_t n = std::min(sizeof(m.msg) - 1, s.size() - sep - 1);
        memcpy(m.msg, s.data() + sep + 1, n);
      }
      out.push_back(m);
    }
  return out;
}

UB_VISIBLE int uccl_engine_send_notif(uccl_conn_t* conn, notify_msg_t* notify_msg) {
  if (!conn || !notify_msg) return -1;
  std::string s(notify_msg->name);
  s.push_back('\0');
  s += std::string(notify_msg->msg);
  return conn->engine->ep->send_notif(conn->id, s) ? 0 : -1;
}

UB_VISIBLE int uccl_engine_prepare_fifo(uccl_engine_t* engine, uccl_mr_t, void const* data, size_t size,
                                        char* fifo_buf) {
  if (!engine || !fifo_buf) return -1;
  XferDesc d;
  if (!engine->ep->describe(data, size, &d)) return -1;
  memcpy(fifo_buf, &d, sizeof(d));
  return 0;
}

// Re-target a descriptor to a sub-window of the same allocation (NIXL does this per request).
UB_VISIBLE int uccl_engine_update_fifo(char* fifo_buf, uint64_t remote_addr, uint32_t size) {
  if (!fifo_buf) return -1;
  XferDesc d;
  memcpy(&d, fifo_buf, sizeof(d));
  d.addr = remote_addr;
  d.size = size;
  memcpy(fifo_buf, &d, sizeof(d));
  return 0;
}

UB_VISIBLE bool uccl_engine_conn_is_local(uccl_conn_t*) { return true; }

UB_VISIBLE int uccl_engine_get_ipc_info(uccl_engine_t* engine, uintptr_t addr, char* ipc_buf, bool* has_ipc) {
  if (!engine || !ipc_buf) return -1;
  XferDesc d;
  bool ok = engine->ep->describe((const void*)addr, 1, &d);
  if (has_ipc) *has_ipc = ok;
  if (!ok) return -1;
  memcpy(ipc_buf, &d, sizeof(d));
  return 0;
}

UB_VISIBLE int uccl_engine_update_ipc_info(char* ipc_buf, uintptr_t addr, size_t size) {
  if (!ipc_buf) return -1;
  XferDesc d;
  memcpy(&d, ipc_buf, sizeof(d));
  d.addr = addr;
  d.size = size;
  memcpy(ipc_buf, &d, sizeof(d));
  return 0;
}
