// Python bindings of the P2P engine (role of the reference's nanobind module p2p/engine_api.cc).
#include <algorithm>

#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "../common/log.h"
#include "compress.h"
#include "endpoint.h"

namespace py = pybind11;
using namespace ub;

namespace {
XferDesc desc_from_bytes(const std::string& s) {
  UB_CHECK(s.size() == sizeof(XferDesc), "descriptor must be %zu bytes (got %zu)", sizeof(XferDesc), s.size());
  XferDesc d;
  memcpy(&d, s.data(), sizeof(d));
  return d;
}
py::bytes desc_to_bytes(const XferDesc& d) { return py::bytes((const char*)&d, sizeof(d)); }
template <typename T>
std::vector<T> to_ptrs(const std::vector<uintptr_t>& v) {
  std::vector<T> o;
  for (auto p : v) o.push_back((T)p);
  return o;
}
}  // namespace

void bind_p2p(py::module_& m) {
  m.attr("XFER_DESC_BYTES") = (int)sizeof(XferDesc);
  py::class_<Endpoint, std::shared_ptr<Endpoint>>(m, "P2PEndpoint")
      .def(py::init<int, int>(), py::arg("local_gpu_idx"), py::arg("num_streams") = 4)
      .def_property_readonly("gpu_idx", &Endpoint::gpu_idx)
      .def_property_readonly("port", &Endpoint::port)
      .def("get_metadata", [](const Endpoint& e) { return py::bytes(e.get_metadata()); })
      .def_static("parse_metadata",
                  [](const std::string& md) {
                    std::string ip;
                    uint16_t port = 0;
                    int gpu = 0;
                    UB_CHECK(Endpoint::parse_metadata(md, &ip, &port, &gpu), "bad metadata blob");
                    return py::make_tuple(ip, (int)port, gpu);
                  })
      .def("connect",
           [](Endpoint& e, const std::string& ip, int gpu, int port) {
             uint64_t id = 0;
             bool ok;
             {
               py::gil_scoped_release rel;
               ok = e.connect(ip, gpu, (uint16_t)port, &id);
             }
             return py::make_tuple(ok, id);
           })
      .def("accept",
           [](Endpoint& e, int timeout_ms) {
             std::string ip;
             int gpu = -1;
             uint64_t id = 0;
             // sliced so that Ctrl-C interrupts a blocking accept (reference: check_python_signals()
             // polled inside blocking loops, p2p/engine.cc:388-392)
             bool ok = false;
             int waited = 0;
             while (!ok && (timeout_ms < 0 || waited < timeout_ms)) {
               const int slice = timeout_ms < 0 ? 100 : std::min(100, timeout_ms - waited);
               {
                 py::gil_scoped_release rel;
                 ok = e.accept(&ip, &gpu, &id, slice);
               }
               waited += slice;
               if (PyErr_CheckSignals() != 0) throw py::error_already_set();
             }
             return py::make_tuple(ok, ip, gpu, id);
           },
           py::arg("timeout_ms") = -1)
      .def("add_remote_endpoint",
           [](Endpoint& e, const std::string& md) {
             uint64_t id = 0;
             bool ok;
             {
               py::gil_scoped_release rel;
               ok = e.add_remote_endpoint(md, &id);
             }
             return py::make_tuple(ok, id);
           })
      .def("remove_remote_endpoint", &Endpoint::remove_remote_endpoint)
      .def("start_passive_accept", &Endpoint::start_passive_accept)
      .def("reg",
           [](Endpoint& e, uintptr_t p, size_t n) {
             uint64_t id = 0;
             bool ok = e.reg((const void*)p, n, &id);
             return py::make_tuple(ok, id);
           })
      .def("dereg", &Endpoint::dereg)
      .def("describe",
           [](Endpoint& e, uintptr_t p, size_t n) {
             XferDesc d;
             UB_CHECK(e.describe((const void*)p, n, &d), "cannot describe pointer %p", (void*)p);
             return desc_to_bytes(d);
           })
      .def("send_async",
           [](Endpoint& e, uint64_t conn, std::vector<uintptr_t> ptrs, std::vector<size_t> sizes) {
             uint64_t tid = 0;
             bool ok = e.send_async(conn, to_ptrs<const void*>(ptrs), sizes, &tid);
             return py::make_tuple(ok, tid);
           })
      .def("recv_async",
           [](Endpoint& e, uint64_t conn, std::vector<uintptr_t> ptrs, std::vector<size_t> sizes) {
             uint64_t tid = 0;
             bool ok = e.recv_async(conn, to_ptrs<void*>(ptrs), sizes, &tid);
             return py::make_tuple(ok, tid);
           })
      .def("prepare",
           [](Endpoint& e, uint64_t conn, bool is_write, std::vector<uintptr_t> ptrs, std::vector<size_t> sizes,
              std::vector<py::bytes> blobs) {
             std::vector<const void*> lp;
             for (auto p : ptrs) lp.push_back((const void*)p);
             std::vector<XferDesc> rd(blobs.size());
             for (size_t i = 0; i < blobs.size(); ++i) {
               std::string b = blobs[i];
               UB_CHECK(b.size() == sizeof(XferDesc), "prepare: descriptor must be %zu bytes", sizeof(XferDesc));
               memcpy(&rd[i], b.data(), sizeof(XferDesc));
             }
             uint64_t id = 0;
             bool ok = e.prepare(conn, is_write, lp, sizes, rd, &id);
             return py::make_tuple(ok, id);
           })
      .def("post",
           [](Endpoint& e, uint64_t prep) {
             uint64_t tid = 0;
             bool ok = e.post(prep, &tid);
             return py::make_tuple(ok, tid);
           })
      .def("release", &Endpoint::release)
      .def("write_async",
           [](Endpoint& e, uint64_t conn, std::vector<uintptr_t> src, std::vector<size_t> sizes,
              std::vector<std::string> remote) {
             std::vector<XferDesc> r;
             for (auto& s : remote) r.push_back(desc_from_bytes(s));
             uint64_t tid = 0;
             bool ok = e.write_async(conn, to_ptrs<const void*>(src), sizes, r, &tid);
             return py::make_tuple(ok, tid);
           })
      .def("read_async",
           [](Endpoint& e, uint64_t conn, std::vector<uintptr_t> dst, std::vector<size_t> sizes,
              std::vector<std::string> remote) {
             std::vector<XferDesc> r;
             for (auto& s : remote) r.push_back(desc_from_bytes(s));
             uint64_t tid = 0;
             bool ok = e.read_async(conn, to_ptrs<void*>(dst), sizes, r, &tid);
             return py::make_tuple(ok, tid);
           })
      .def("poll_async",
           [](Endpoint& e, uint64_t tid) {
             bool done = false;
             bool ok = e.poll_async(tid, &done);
             return py::make_tuple(ok, done);
           })
      .def("wait",
           [](Endpoint& e, uint64_t tid, int timeout_ms) {
             bool done = false;
             int waited = 0;
             while (!done && (timeout_ms < 0 || waited < timeout_ms)) {
               const int slice = timeout_ms < 0 ? 100 : std::min(100, timeout_ms - waited);
               bool valid = true;
               {
                 py::gil_scoped_release rel;
                 done = e.wait(tid, slice);
                 if (!done) valid = e.poll_async(tid, &done);  // false: unknown / failed transfer, not a timeout
               }
               if (!valid) return false;
               waited += slice;
               if (!done && PyErr_CheckSignals() != 0) throw py::error_already_set();
             }
             return done;
           },
           py::arg("tid"), py::arg("timeout_ms") = -1)
      .def("send_notif", &Endpoint::send_notif)
      .def("get_notifs",
           [](Endpoint& e) {
             py::list out;
             for (auto& kv : e.get_notifs()) out.append(py::make_tuple(kv.first, py::bytes(kv.second)));
             return out;
           })
      .def("stats", [](const Endpoint& e) {
        P2PStats s = e.stats();
        py::dict d;
        d["bytes_sent"] = s.bytes_sent;
        d["bytes_received"] = s.bytes_received;
        d["bytes_read"] = s.bytes_read;
        d["bytes_written"] = s.bytes_written;
        d["transfers"] = s.transfers;
        d["kernel_launches"] = s.kernel_launches;
        d["memcpy_fallbacks"] = s.memcpy_fallbacks;
        return d;
      });

  // ---- compression hook
  m.attr("CMP_HEADER_BYTES") = (int)sizeof(CmpHeader);
  m.def("cmp_supported", &cmp_dtype_supported);
  m.def("cmp_bound", &cmp_bound);
  m.def("cmp_compress", [](uintptr_t src, size_t count, int dtype, uintptr_t dst, uintptr_t st) {
    cudaError_t e = cmp_compress_async((const void*)src, count, dtype, (void*)dst, (cudaStream_t)st);
    UB_CHECK(e == cudaSuccess, "cmp_compress: %s", cudaGetErrorString(e));
  });
  m.def("cmp_decompress", [](uintptr_t src, uintptr_t dst, size_t count, int dtype, uintptr_t st) {
    cudaError_t e = cmp_decompress_async((const void*)src, (void*)dst, count, dtype, (cudaStream_t)st);
    UB_CHECK(e == cudaSuccess, "cmp_decompress: %s", cudaGetErrorString(e));
  });
}
