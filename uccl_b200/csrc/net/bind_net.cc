// Python surface of the inter-node datagram transport: `uccl_b200._C.net`.
// Buffers cross as integer addresses (tensor.data_ptr() of CPU / pinned tensors); requests as opaque ints.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <chrono>
#include <thread>

#include "../common/log.h"
#include "../common/timers.h"
#include "net_engine.h"

namespace py = pybind11;
using namespace ub::net;

void bind_net(py::module_& root) {
  py::module_ m = root.def_submodule("net", "multipath reliable datagram transport between nodes");
  m.attr("CC_NONE") = (int)CC_NONE;
  m.attr("CC_SWIFT") = (int)CC_SWIFT;
  m.attr("CC_TIMELY") = (int)CC_TIMELY;
  m.attr("CC_EQDS") = (int)CC_EQDS;
  m.attr("FL_SYN_SENT") = (int)FL_SYN_SENT;
  m.attr("FL_ESTABLISHED") = (int)FL_ESTABLISHED;
  m.attr("FL_CLOSING") = (int)FL_CLOSING;
  m.attr("FL_CLOSED") = (int)FL_CLOSED;
  m.attr("FL_ERROR") = (int)FL_ERROR;
  m.attr("MAX_PATHS") = kMaxPaths;
  m.attr("HEADER_BYTES") = (int)sizeof(PktHdr);
  m.def("list_interfaces", &list_interfaces);

  py::class_<Engine, std::shared_ptr<Engine>>(m, "Engine")
      .def(py::init([](const std::string& bind_ip, int paths, int payload, int max_inflight, long eager_max, int cc,
                       double drop_prob, int rto_min_us, int rto_abort, double link_gbps, bool busy_poll) {
             EngineConfig c = EngineConfig::from_env();
             if (!bind_ip.empty()) c.bind_ip = bind_ip;
             if (paths > 0) c.paths = paths;
             if (payload > 0) c.payload = payload;
             if (max_inflight > 0) c.max_inflight = max_inflight;
             if (eager_max >= 0) c.eager_max = (size_t)eager_max;
             if (cc >= 0) c.cc = cc;
             if (drop_prob >= 0) c.drop_prob = drop_prob;
             if (rto_min_us > 0) c.rto_min_us = rto_min_us;
             if (rto_abort > 0) c.rto_abort = rto_abort;
             if (link_gbps > 0) c.link_gbps = link_gbps;
             c.busy_poll = c.busy_poll || busy_poll;
             // the destructor lingers for the FIN exchange: never do that while holding the GIL
             return std::shared_ptr<Engine>(new Engine(c), [](Engine* e) {
               // the last reference may be dropped by Python (GIL held) or by a native owner that already
               // released it (e.g. a ukernel communicator being destroyed)
               if (PyGILState_Check()) {
                 py::gil_scoped_release rel;
                 delete e;
               } else {
                 delete e;
               }
             });
           }),
           py::arg("bind_ip") = "", py::arg("paths") = 0, py::arg("payload") = 0, py::arg("max_inflight") = 0,
           py::arg("eager_max") = -1, py::arg("cc") = -1, py::arg("drop_prob") = -1.0, py::arg("rto_min_us") = 0,
           py::arg("rto_abort") = 0, py::arg("link_gbps") = 0.0, py::arg("busy_poll") = false)
      .def_property_readonly("port", &Engine::port)
      .def_property_readonly("paths", &Engine::paths)
      .def_property_readonly("bind_ip", [](Engine& e) { return e.config().bind_ip; })
      .def_property_readonly("payload", [](Engine& e) { return e.config().payload; })
      .def("listen", &Engine::listen)
      .def("close_listen", &Engine::close_listen)
      .def("connect_async", &Engine::connect_async)
      .def("flow_state", &Engine::flow_state)
      .def("accept_nb",
           [](Engine& e, uint32_t lid) -> py::object {
             uint32_t f = 0;
             if (!e.accept_nb(lid, &f)) return py::none();
             return py::int_(f);
           })
      .def(
          "connect",
          [](Engine& e, const std::string& ip, uint16_t port, uint32_t lid, int timeout_ms) {
            py::gil_scoped_release rel;
            return e.connect(ip, port, lid, timeout_ms);
          },
          py::arg("ip"), py::arg("port"), py::arg("listen_id"), py::arg("timeout_ms") = 30000)
      .def(
          "accept",
          [](Engine& e, uint32_t lid, int timeout_ms) {
            py::gil_scoped_release rel;
            return e.accept(lid, timeout_ms);
          },
          py::arg("listen_id"), py::arg("timeout_ms") = 30000)
      .def("close_flow", &Engine::close_flow)
      .def("shutdown", &Engine::shutdown, py::arg("linger_ms") = -1, py::call_guard<py::gil_scoped_release>())
      .def("send_async",
           [](Engine& e, uint32_t flow, uintptr_t ptr, size_t n) { return (uintptr_t)e.send_async(flow, (const void*)ptr, n); })
      .def("recv_async",
           [](Engine& e, uint32_t flow, uintptr_t ptr, size_t cap) { return (uintptr_t)e.recv_async(flow, (void*)ptr, cap); })
      .def("test",
           [](Engine& e, uintptr_t req) -> py::object {
             size_t bytes = 0;
             int err = 0;
             if (!e.test((Request*)req, &bytes, &err)) return py::none();
             return py::make_tuple(bytes, err);
           })
      .def(
          "wait",
          [](Engine& e, uintptr_t req, int timeout_ms) {
            size_t bytes = 0;
            int err = 0;
            bool done = false;
            {
              py::gil_scoped_release rel;
              const uint64_t t0 = ub::now_ns();
              uint32_t spins = 0;
              while (!(done = e.test((Request*)req, &bytes, &err))) {
                if (timeout_ms >= 0 && ub::now_ns() - t0 > (uint64_t)timeout_ms * 1000000ull) break;
                if (++spins > 2000) std::this_thread::sleep_for(std::chrono::microseconds(20));
                else std::this_thread::yield();
              }
            }
            UB_CHECK(done, "net: request timed out after %d ms", timeout_ms);
            UB_CHECK(err == 0, "net: request failed: %s",
                     err == 2 ? "message larger than the posted receive" : err == 3 ? "peer closed the flow" : "flow error");
            return bytes;
          },
          py::arg("request"), py::arg("timeout_ms") = -1)
      .def("set_drop_prob", &Engine::set_drop_prob)
      .def("set_reorder", &Engine::set_reorder, py::arg("prob"), py::arg("delay_us") = 300)
      .def("set_path_drop", &Engine::set_path_drop, py::arg("path"), py::arg("prob") = 1.0)
      .def("stats",
           [](Engine& e) {
             const EngineStats s = e.stats();
             py::dict d;
             d["tx_pkts"] = s.tx_pkts, d["rx_pkts"] = s.rx_pkts, d["tx_bytes"] = s.tx_bytes, d["rx_bytes"] = s.rx_bytes;
             d["dropped_tx"] = s.dropped_tx, d["bad_pkts"] = s.bad_pkts, d["fast_rexmit"] = s.fast_rexmit;
             d["rto_rexmit"] = s.rto_rexmit, d["loops"] = s.loops, d["sleeps"] = s.sleeps, d["flows"] = s.flows;
             return d;
           })
      .def("flow_stats", [](Engine& e, uint32_t flow) -> py::object {
        FlowStats s;
        if (!e.flow_stats(flow, &s)) return py::none();
        py::dict d;
        d["tx_pkts"] = s.tx_pkts, d["tx_bytes"] = s.tx_bytes, d["rx_pkts"] = s.rx_pkts, d["rx_bytes"] = s.rx_bytes;
        d["rx_dup"] = s.rx_dup, d["fast_rexmit"] = s.fast_rexmit, d["rto_rexmit"] = s.rto_rexmit, d["tlp"] = s.tlp;
        d["acks_tx"] = s.acks_tx, d["acks_rx"] = s.acks_rx, d["unexpected_msgs"] = s.unexpected_msgs;
        d["srtt_us"] = s.srtt_us, d["min_rtt_us"] = s.min_rtt_us, d["cwnd"] = s.cwnd, d["rate_gbps"] = s.rate_gbps;
        d["state"] = s.state;
        d["path_bans"] = s.path_bans;
        std::vector<uint64_t> paths(s.path_tx, s.path_tx + e.paths());
        d["path_tx"] = paths;
        return d;
      });
}
