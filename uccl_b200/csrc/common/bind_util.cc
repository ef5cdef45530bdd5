// Python bindings of the host utilities (for tests and for tooling built on them).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <thread>

#include "cc/eqds.h"
#include "cc/swift.h"
#include "cc/timely.h"
#include "latency.h"
#include "log.h"
#include "param.h"
#include "pool.h"
#include "ring.h"
#include "sm_partition.h"
#include "../kernels/launch.h"
#include "timers.h"
#include "timing_wheel.h"

namespace py = pybind11;
using namespace ub;

void bind_util(py::module_& m) {
  py::module_ u = m.def_submodule("util", "host utilities: lock-free rings, pools, histograms, congestion control");
  u.def("now_ns", &now_ns);
  u.def("tsc_ghz", &tsc_ghz);
  u.def("log", [](int level, const std::string& msg) { log_emit(level, SUB_UTIL, "python", 0, "%s", msg.c_str()); });
  // SM partitions (CUDA green contexts): see sm_partition.h
  py::class_<SmPartition, std::shared_ptr<SmPartition>>(u, "SmPartition")
      .def_static("supported",
                  [](int device) {
                    std::string why;
                    const bool ok = SmPartition::supported(device, &why);
                    return py::make_tuple(ok, why);
                  })
      .def_static("device_sm_count", &SmPartition::device_sm_count)
      .def_static(
          "split",
          [](int device, int sm_count, bool fine_grained) {
            SmPartition::Pair p = SmPartition::split(device, sm_count, fine_grained);
            return py::make_tuple(p.part, p.rest);
          },
          py::arg("device"), py::arg("sm_count"), py::arg("fine_grained") = false)
      .def_property_readonly("device", &SmPartition::device)
      .def_property_readonly("sm_count", &SmPartition::sm_count)
      .def("stream", [](SmPartition& p, int priority) { return (uintptr_t)p.stream(priority); }, py::arg("priority") = 0);

  u.def(
      "smid_probe",
      [](uintptr_t out, int blocks, uint64_t hold_ns, uintptr_t stream) {
        cudaError_t e = launch_smid_probe(reinterpret_cast<int*>(out), blocks, hold_ns, reinterpret_cast<cudaStream_t>(stream));
        UB_CHECK(e == cudaSuccess, "smid probe launch failed: %s", cudaGetErrorString(e));
      },
      py::arg("out"), py::arg("blocks"), py::arg("hold_ns") = 20000, py::arg("stream") = 0,
      "out[b] (int32, device) = id of the SM CTA b ran on");

  u.def("param_str", [](const std::string& k, const std::string& d) { return param_load_str(k.c_str(), d.c_str()); });

  py::class_<SpscRing<uint64_t>>(u, "SpscRing")
      .def(py::init<size_t>())
      .def("push", &SpscRing<uint64_t>::push)
      .def("pop",
           [](SpscRing<uint64_t>& r) -> py::object {
             uint64_t v;
             if (r.pop(&v)) return py::int_(v);
             return py::none();
           })
      .def("size", &SpscRing<uint64_t>::size)
      .def_property_readonly("capacity", &SpscRing<uint64_t>::capacity);

  py::class_<MpmcRing<uint64_t>>(u, "MpmcRing")
      .def(py::init<size_t>())
      .def("push", &MpmcRing<uint64_t>::push)
      .def("pop",
           [](MpmcRing<uint64_t>& r) -> py::object {
             uint64_t v;
             if (r.pop(&v)) return py::int_(v);
             return py::none();
           })
      .def_property_readonly("capacity", &MpmcRing<uint64_t>::capacity)
      // native multi-threaded stress: P producers push [0, per) tagged values, C consumers drain;
      // returns (sum popped, count popped)
      .def("stress",
           [](MpmcRing<uint64_t>& r, int producers, int consumers, uint64_t per) {
             py::gil_scoped_release rel;
             std::atomic<uint64_t> sum{0}, cnt{0};
             std::atomic<int> live{producers};
             std::vector<std::thread> ts;
             for (int p = 0; p < producers; ++p)
               ts.emplace_back([&, p] {
                 for (uint64_t i = 0; i < per; ++i) {
                   const uint64_t v = ((uint64_t)p << 40) | (i + 1);
                   while (!r.push(v)) std::this_thread::yield();
                 }
                 live.fetch_sub(1);
               });
             for (int c = 0; c < consumers; ++c)
               ts.emplace_back([&] {
                 uint64_t v;
                 for (;;) {
                   if (r.pop(&v)) {
                     sum.fetch_add(v & ((1ull << 40) - 1));
                     cnt.fetch_add(1);
                   } else if (live.load() == 0) {
                     if (!r.pop(&v)) break;
                     sum.fetch_add(v & ((1ull << 40) - 1));
                     cnt.fetch_add(1);
                   } else {
                     std::this_thread::yield();
                   }
                 }
               });
             for (auto& t : ts) t.join();
             return std::make_pair(sum.load(), cnt.load());
           });

  py::class_<SharedPool<uint64_t>>(u, "SharedPool")
      .def(py::init<size_t>())
      .def("release_global", &SharedPool<uint64_t>::release_global)
      .def("put", &SharedPool<uint64_t>::put)
      .def("get",
           [](SharedPool<uint64_t>& p) -> py::object {
             uint64_t v;
             if (p.get(&v)) return py::int_(v);
             return py::none();
           })
      .def("global_size", &SharedPool<uint64_t>::global_size);

  py::class_<LatencyHist>(u, "LatencyHist")
      .def(py::init<>())
      .def("record", &LatencyHist::record)
      .def("count", &LatencyHist::count)
      .def("mean", &LatencyHist::mean)
      .def("min", &LatencyHist::min)
      .def("max", &LatencyHist::max)
      .def("percentile", &LatencyHist::percentile)
      .def("merge", &LatencyHist::merge)
      .def("reset", &LatencyHist::reset)
      .def("summary", &LatencyHist::summary, py::arg("unit") = "ns");

  u.def("seqno_less", [](unsigned bits, uint32_t a, uint32_t b) {
    if (bits == 16) return SeqNo<16>(a) < SeqNo<16>(b);
    if (bits == 8) return SeqNo<8>(a) < SeqNo<8>(b);
    return SeqNo<32>(a) < SeqNo<32>(b);
  });

  py::class_<TimingWheel<uint64_t>>(u, "TimingWheel")
      .def(py::init<uint64_t, size_t, uint64_t>(), py::arg("granularity_ns"), py::arg("slots"), py::arg("now_ns") = 0)
      .def("insert", &TimingWheel<uint64_t>::insert)
      .def("advance",
           [](TimingWheel<uint64_t>& w, uint64_t now) {
             std::vector<uint64_t> out;
             w.advance(now, &out);
             return out;
           })
      .def("__len__", &TimingWheel<uint64_t>::size)
      .def_property_readonly("horizon_ns", &TimingWheel<uint64_t>::horizon_ns);
  py::class_<cc::Timely>(u, "Timely")
      .def(py::init<>())
      .def("on_rtt", &cc::Timely::on_rtt)
      .def("rate_gbps", &cc::Timely::rate_gbps)
      .def("pacing_delay_us", &cc::Timely::pacing_delay_us);
  py::class_<cc::Swift>(u, "Swift")
      .def(py::init<>())
      .def("on_ack", &cc::Swift::on_ack, py::arg("delay_us"), py::arg("acked"), py::arg("now_us"), py::arg("rtt_us"),
           py::arg("hops") = 1)
      .def("cwnd", &cc::Swift::cwnd)
      .def("target_delay_us", &cc::Swift::target_delay_us, py::arg("hops") = 1)
      .def("on_retransmit_timeout", &cc::Swift::on_retransmit_timeout)
      .def("pacing_delay_us", &cc::Swift::pacing_delay_us);
  py::class_<cc::EqdsPacer>(u, "EqdsPacer")
      .def(py::init<>())
      .def("add_demand", &cc::EqdsPacer::add_demand)
      .def("on_data", &cc::EqdsPacer::on_data)
      .def("tick", &cc::EqdsPacer::tick)
      .def("granted", &cc::EqdsPacer::granted)
      .def("active_senders", &cc::EqdsPacer::active_senders);
}
