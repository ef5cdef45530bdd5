// Python bindings of the ukernel subsystem (worker FIFOs, planner, communicator).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "../common/log.h"
#include "uk_comm.h"
#include "uk_net.h"

namespace py = pybind11;
using namespace ub;

namespace {
py::list plan_ops(const UkPlan& p) {
  py::list out;
  for (const auto& o : p.ops) {
    py::dict d;
    static const char* kinds[] = {"copy", "reduce", "send", "recv"};
    static const char* bufs[] = {"in", "out", "scratch"};
    d["kind"] = kinds[o.kind];
    d["lane"] = o.lane;
    d["tile"] = o.tile;
    d["step"] = o.step;
    d["peer"] = o.peer;
    d["bytes"] = o.bytes;
    d["dst"] = py::make_tuple(bufs[(int)o.dst.buf], o.dst.off);
    d["src"] = py::make_tuple(bufs[(int)o.src.buf], o.src.off);
    d["src2"] = py::make_tuple(bufs[(int)o.src2.buf], o.src2.off);
    d["deps"] = o.deps;
    out.append(d);
  }
  return out;
}
// inverse of plan_ops(): a user-authored plan (uccl_b200.ukernel.dsl) -> UkPlan
UkRef ref_from(const py::handle& h) {
  py::tuple t = py::reinterpret_borrow<py::tuple>(h);
  UB_CHECK(t.size() == 2, "ukernel plan: a buffer reference is (name, offset)");
  const std::string name = t[0].cast<std::string>();
  UkRef r;
  r.buf = name == "in" ? UkBuf::In : (name == "out" ? UkBuf::Out : UkBuf::Scratch);
  UB_CHECK(name == "in" || name == "out" || name == "scratch", "ukernel plan: unknown buffer '%s'", name.c_str());
  r.off = t[1].cast<uint64_t>();
  return r;
}
UkPlan plan_from(int nranks, int rank, int nlanes, uint64_t scratch_bytes, const py::list& ops) {
  UkPlan p;
  p.coll = UkColl::AllReduce;  // informational only for custom programs
  p.algo = UkAlgo::Auto;
  p.nranks = nranks, p.rank = rank, p.nlanes = nlanes, p.scratch_bytes = scratch_bytes;
  static const char* kinds[] = {"copy", "reduce", "send", "recv"};
  for (auto h : ops) {
    py::dict d = py::reinterpret_borrow<py::dict>(h);
    UkPlanOp o;
    const std::string k = d["kind"].cast<std::string>();
    o.kind = -1;
    for (int i = 0; i < 4; ++i)
      if (k == kinds[i]) o.kind = i;
    UB_CHECK(o.kind >= 0, "ukernel plan: unknown op kind '%s'", k.c_str());
    o.lane = d.contains("lane") ? d["lane"].cast<int>() : 0;
    o.tile = d.contains("tile") ? d["tile"].cast<int>() : 0;
    o.step = d.contains("step") ? d["step"].cast<int>() : 0;
    o.peer = d.contains("peer") ? d["peer"].cast<int>() : -1;
    o.bytes = d.contains("bytes") ? d["bytes"].cast<uint64_t>() : 0;
    if (d.contains("dst")) o.dst = ref_from(d["dst"]);
    if (d.contains("src")) o.src = ref_from(d["src"]);
    if (d.contains("src2")) o.src2 = ref_from(d["src2"]);
    if (d.contains("deps")) o.deps = d["deps"].cast<std::vector<int>>();
    p.ops.push_back(o);
  }
  return p;
}
std::vector<UkPlan> plans_from(int nlanes, const std::vector<uint64_t>& scratch, const py::list& per_rank) {
  const int n = (int)per_rank.size();
  UB_CHECK((int)scratch.size() == n, "ukernel plans: one scratch size per rank");
  std::vector<UkPlan> plans;
  for (int r = 0; r < n; ++r) plans.push_back(plan_from(n, r, nlanes, scratch[r], py::reinterpret_borrow<py::list>(per_rank[r])));
  return plans;
}
UkPlan make(int coll, uint64_t bytes, int nranks, int rank, int nlanes, uint64_t tile, uint64_t elem, int algo,
            int root = 0) {
  UkPlanParams p;
  p.nranks = nranks, p.rank = rank, p.nlanes = nlanes, p.tile_bytes = tile, p.elem_size = elem, p.algo = (UkAlgo)algo;
  switch ((UkColl)coll) {
    case UkColl::AllReduce: return uk_plan_allreduce(bytes, p);
    case UkColl::AllToAll: return uk_plan_alltoall(bytes, p);
    case UkColl::AllGather: return uk_plan_allgather(bytes, p);
    case UkColl::ReduceScatter: return uk_plan_reduce_scatter(bytes, p);
    case UkColl::Broadcast: return uk_plan_broadcast(bytes, root, p);
    default: return uk_plan_barrier(p);
  }
}
}  // namespace

void bind_uk(py::module_& m) {
  py::module_ uk = m.def_submodule("uk", "persistent-worker task executor + CCL planner");
  uk.attr("ALLREDUCE") = (int)UkColl::AllReduce;
  uk.attr("ALLTOALL") = (int)UkColl::AllToAll;
  uk.attr("ALLGATHER") = (int)UkColl::AllGather;
  uk.attr("BARRIER") = (int)UkColl::Barrier;
  uk.attr("REDUCE_SCATTER") = (int)UkColl::ReduceScatter;
  uk.attr("BROADCAST") = (int)UkColl::Broadcast;
  uk.attr("ALGO_AUTO") = (int)UkAlgo::Auto;
  uk.attr("ALGO_RING") = (int)UkAlgo::Ring;
  uk.attr("ALGO_FULLMESH") = (int)UkAlgo::FullMesh;
  uk.attr("OP_COPY") = (int)UK_COPY;
  uk.attr("OP_REDUCE") = (int)UK_REDUCE;
  uk.attr("OP_SIGNAL") = (int)UK_SIGNAL;
  uk.attr("OP_WAIT") = (int)UK_WAIT;
  uk.attr("RING_ENTRIES") = (int)kUkRingEntries;

  uk.def("select_algo", [](int coll, int nranks, uint64_t bytes) { return (int)uk_select_algo((UkColl)coll, nranks, bytes); });
  uk.def("scratch_bytes", [](int algo, int nranks, int nlanes, uint64_t tile) {
    return uk_scratch_bytes((UkAlgo)algo, nranks, nlanes, tile);
  });
  // plan for one rank: (description, [op dicts])
  uk.def("plan",
         [](int coll, uint64_t bytes, int nranks, int rank, int nlanes, uint64_t tile, uint64_t elem, int algo, int root) {
           UkPlan p = make(coll, bytes, nranks, rank, nlanes, tile, elem, algo, root);
           return py::make_tuple(p.describe(), plan_ops(p));
         },
         py::arg("coll"), py::arg("bytes"), py::arg("nranks"), py::arg("rank"), py::arg("nlanes") = 1,
         py::arg("tile_bytes") = 1 << 20, py::arg("elem_size") = 1, py::arg("algo") = 0, py::arg("root") = 0);
  // plans of all ranks: structural validation, "" when consistent
  uk.def("validate",
         [](int coll, uint64_t bytes, int nranks, int nlanes, uint64_t tile, uint64_t elem, int algo, int root) {
           std::vector<UkPlan> plans;
           for (int r = 0; r < nranks; ++r) plans.push_back(make(coll, bytes, nranks, r, nlanes, tile, elem, algo, root));
           return uk_validate(plans);
         },
         py::arg("coll"), py::arg("bytes"), py::arg("nranks"), py::arg("nlanes") = 1, py::arg("tile_bytes") = 1 << 20,
         py::arg("elem_size") = 1, py::arg("algo") = 0, py::arg("root") = 0);
  // run the plans of all ranks over host buffers (in_ptrs / out_ptrs: one per rank); "" on success
  uk.def("simulate",
         [](int coll, uint64_t bytes, int nranks, int nlanes, uint64_t tile, int dtype, int op, int algo,
            std::vector<uintptr_t> in_ptrs, std::vector<uintptr_t> out_ptrs, int root) {
           UB_CHECK((int)in_ptrs.size() == nranks && (int)out_ptrs.size() == nranks, "simulate: one buffer per rank");
           std::vector<UkPlan> plans;
           for (int r = 0; r < nranks; ++r)
             plans.push_back(make(coll, bytes, nranks, r, nlanes, tile, dtype_size(dtype), algo, root));
           std::string err = uk_validate(plans);
           if (!err.empty()) return err;
           std::vector<std::vector<char>> scratch(nranks);
           UkSimBuffers b;
           for (int r = 0; r < nranks; ++r) {
             scratch[r].assign(plans[r].scratch_bytes + 16, 0x5a);
             b.in.push_back((char*)in_ptrs[r]);
             b.out.push_back((char*)out_ptrs[r]);
             b.scratch.push_back(scratch[r].data());
           }
           py::gil_scoped_release rel;
           return uk_simulate(plans, b, dtype, op);
         },
         py::arg("coll"), py::arg("bytes"), py::arg("nranks"), py::arg("nlanes"), py::arg("tile_bytes"), py::arg("dtype"),
         py::arg("op"), py::arg("algo"), py::arg("in_ptrs"), py::arg("out_ptrs"), py::arg("root") = 0);

  // ---- user-authored programs (uccl_b200.ukernel.dsl): per-rank op lists in the format uk.plan() returns
  uk.def("validate_ops",
         [](int nlanes, std::vector<uint64_t> scratch, py::list per_rank, uint64_t in_bytes, uint64_t out_bytes,
            uint64_t elem_size) {
           std::vector<UkPlan> plans = plans_from(nlanes, scratch, per_rank);
           for (auto& p : plans) {
             std::string e = uk_check_bounds(p, in_bytes, out_bytes, p.scratch_bytes, nlanes, elem_size);
             if (!e.empty()) return e;
           }
           return uk_validate(plans);
         },
         py::arg("nlanes"), py::arg("scratch_bytes"), py::arg("ops"), py::arg("in_bytes"), py::arg("out_bytes"),
         py::arg("elem_size") = 1);
  uk.def("simulate_ops",
         [](int nlanes, std::vector<uint64_t> scratch, py::list per_rank, uint64_t in_bytes, uint64_t out_bytes, int dtype,
            int op, std::vector<uintptr_t> in_ptrs, std::vector<uintptr_t> out_ptrs) {
           std::vector<UkPlan> plans = plans_from(nlanes, scratch, per_rank);
           const int n = (int)plans.size();
           UB_CHECK((int)in_ptrs.size() == n && (int)out_ptrs.size() == n, "simulate_ops: one buffer per rank");
           for (auto& p : plans) {
             std::string e = uk_check_bounds(p, in_bytes, out_bytes, p.scratch_bytes, nlanes, dtype_size(dtype));
             if (!e.empty()) return e;
           }
           std::string err = uk_validate(plans);
           if (!err.empty()) return err;
           std::vector<std::vector<char>> sc(n);
           UkSimBuffers b;
           for (int r = 0; r < n; ++r) {
             sc[r].assign(plans[r].scratch_bytes + 16, 0x5a);
             b.in.push_back((char*)in_ptrs[r]);
             b.out.push_back((char*)out_ptrs[r]);
             b.scratch.push_back(sc[r].data());
           }
           py::gil_scoped_release rel;
           return uk_simulate(plans, b, dtype, op);
         },
         py::arg("nlanes"), py::arg("scratch_bytes"), py::arg("ops"), py::arg("in_bytes"), py::arg("out_bytes"),
         py::arg("dtype"), py::arg("op"), py::arg("in_ptrs"), py::arg("out_ptrs"));

  py::class_<UkWorker, std::shared_ptr<UkWorker>>(uk, "Worker")
      .def(py::init<int, int, uint64_t, int64_t>(), py::arg("device"), py::arg("nlanes"), py::arg("timeout_ms") = 20000,
           py::arg("idle_us") = -1)
      .def_property_readonly("kernel_launches", &UkWorker::kernel_launches)
      .def("start", [](UkWorker& w) { w.start(); })
      .def("stop", &UkWorker::stop, py::call_guard<py::gil_scoped_release>())
      .def_property_readonly("running", &UkWorker::running)
      .def_property_readonly("nlanes", &UkWorker::nlanes)
      .def_property_readonly("is_host", &UkWorker::is_host)
      .def("push",
           [](UkWorker& w, int lane, int op, uintptr_t dst, uintptr_t src, uintptr_t src2, uint64_t bytes, int dtype,
              int redop, uintptr_t sig_addr, uint64_t sig_val) {
             UkTask t;
             memset(&t, 0, sizeof(t));
             t.op = (uint32_t)op, t.dtype = (uint32_t)dtype, t.redop = (uint32_t)redop;
             t.src = src, t.src2 = src2;
             if (op == UK_SIGNAL || op == UK_WAIT) t.sig_addr = sig_addr, t.sig_val = sig_val;
             else t.dst = dst, t.bytes = bytes;
             py::gil_scoped_release rel;
             return w.push(lane, t);
           },
           py::arg("lane"), py::arg("op"), py::arg("dst") = 0, py::arg("src") = 0, py::arg("src2") = 0,
           py::arg("bytes") = 0, py::arg("dtype") = 0, py::arg("redop") = 0, py::arg("sig_addr") = 0,
           py::arg("sig_val") = 0)
      .def("completed", &UkWorker::completed)
      .def("done", &UkWorker::done)
      .def("wait", &UkWorker::wait, py::arg("lane"), py::arg("ticket"), py::arg("timeout_s") = 30.0,
           py::call_guard<py::gil_scoped_release>())
      .def("wait_all", &UkWorker::wait_all, py::arg("timeout_s") = 30.0, py::call_guard<py::gil_scoped_release>())
      .def("done_counter", [](UkWorker& w, int lane) { return (uintptr_t)w.done_counter(lane); })
      .def_property_readonly("error", &UkWorker::error)
      .def("stats", [](UkWorker& w) {
        auto s = w.stats();
        py::dict d;
        d["pushed"] = s.pushed, d["copies"] = s.copies, d["reduces"] = s.reduces, d["signals"] = s.signals;
        d["waits"] = s.waits, d["bytes"] = s.bytes;
        return d;
      });

  py::class_<UkComm, std::shared_ptr<UkComm>>(uk, "Comm")
      .def(py::init([](std::shared_ptr<Comm> c, int nlanes, uint64_t tile, uint64_t staging) {
             UkCommConfig cfg;
             cfg.nlanes = nlanes, cfg.tile_bytes = tile, cfg.staging_bytes = staging;
             py::gil_scoped_release rel;
             return std::make_shared<UkComm>(c, cfg);
           }),
           py::arg("comm"), py::arg("nlanes") = 4, py::arg("tile_bytes") = 1 << 20, py::arg("staging_bytes") = 32 << 20)
      .def_property_readonly("rank", &UkComm::rank)
      .def_property_readonly("nranks", &UkComm::nranks)
      .def_property_readonly("nlanes", &UkComm::nlanes)
      .def("all_reduce",
           [](UkComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype, int op, int algo, uintptr_t stream,
              bool symmetric) {
             py::gil_scoped_release rel;
             return u.all_reduce((const void*)in, (void*)out, count, dtype, op, (UkAlgo)algo, (cudaStream_t)stream,
                                 symmetric);
           },
           py::arg("inp"), py::arg("out"), py::arg("count"), py::arg("dtype"), py::arg("op"), py::arg("algo") = 0,
           py::arg("stream") = 0, py::arg("symmetric") = false)
      .def("all_to_all",
           [](UkComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype, uintptr_t stream, bool symmetric) {
             py::gil_scoped_release rel;
             return u.all_to_all((const void*)in, (void*)out, count, dtype, (cudaStream_t)stream, symmetric);
           },
           py::arg("inp"), py::arg("out"), py::arg("count_per_peer"), py::arg("dtype"), py::arg("stream") = 0,
           py::arg("symmetric") = false)
      .def("all_gather",
           [](UkComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype, uintptr_t stream, bool symmetric) {
             py::gil_scoped_release rel;
             return u.all_gather((const void*)in, (void*)out, count, dtype, (cudaStream_t)stream, symmetric);
           },
           py::arg("inp"), py::arg("out"), py::arg("count_per_rank"), py::arg("dtype"), py::arg("stream") = 0,
           py::arg("symmetric") = false)
      .def("reduce_scatter",
           [](UkComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype, int op, uintptr_t stream) {
             py::gil_scoped_release rel;
             return u.reduce_scatter((const void*)in, (void*)out, count, dtype, op, (cudaStream_t)stream);
           },
           py::arg("inp"), py::arg("out"), py::arg("recv_count"), py::arg("dtype"), py::arg("op"), py::arg("stream") = 0)
      .def("broadcast",
           [](UkComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype, int root, uintptr_t stream) {
             py::gil_scoped_release rel;
             return u.broadcast((const void*)in, (void*)out, count, dtype, root, (cudaStream_t)stream);
           },
           py::arg("inp"), py::arg("out"), py::arg("count"), py::arg("dtype"), py::arg("root"), py::arg("stream") = 0)
      .def("barrier",
           [](UkComm& u, uintptr_t stream) {
             py::gil_scoped_release rel;
             return u.barrier((cudaStream_t)stream);
           },
           py::arg("stream") = 0)
      .def("run_custom",
           [](UkComm& u, int nlanes, uint64_t scratch_bytes, py::list ops, uintptr_t in, uint64_t in_bytes, uintptr_t out,
              uint64_t out_bytes, int dtype, int op, uintptr_t stream, bool symmetric) {
             UkPlan p = plan_from(u.nranks(), u.rank(), nlanes, scratch_bytes, ops);
             py::gil_scoped_release rel;
             return u.run_custom(p, (const void*)in, in_bytes, (void*)out, out_bytes, dtype, op, (cudaStream_t)stream, symmetric);
           },
           py::arg("nlanes"), py::arg("scratch_bytes"), py::arg("ops"), py::arg("in_ptr"), py::arg("in_bytes"),
           py::arg("out_ptr"), py::arg("out_bytes"), py::arg("dtype"), py::arg("op"), py::arg("stream") = 0,
           py::arg("symmetric") = false)
      .def("scratch_capacity", &UkComm::scratch_capacity)
      .def("test", &UkComm::test)
      .def("wait", &UkComm::wait, py::arg("ticket"), py::arg("timeout_s") = 60.0, py::call_guard<py::gil_scoped_release>())
      .def("stop", &UkComm::stop, py::call_guard<py::gil_scoped_release>())
      .def_property_readonly("kernel_launches", [](UkComm& u) { return u.worker().kernel_launches(); })
      .def("stats", [](UkComm& u) {
        auto s = u.stats();
        py::dict d;
        d["ops"] = s.ops, d["segments"] = s.segments, d["tasks"] = s.tasks, d["zero_copy_ops"] = s.zero_copy_ops;
        d["stream_ordered_ops"] = s.stream_ordered_ops;
        return d;
      });
  // plans over the message transport (ranks on other boxes)
  py::class_<UkNetComm, std::shared_ptr<UkNetComm>>(uk, "UkNetComm")
      .def(py::init([](int rank, int nranks, std::shared_ptr<net::Engine> engine, std::vector<uint32_t> flows, int nlanes,
                       uint64_t tile_bytes, int timeout_ms) {
             UkNetConfig c;
             c.nlanes = nlanes, c.tile_bytes = tile_bytes, c.timeout_ms = timeout_ms;
             return std::shared_ptr<UkNetComm>(new UkNetComm(rank, nranks, std::move(engine), std::move(flows), c),
                                               [](UkNetComm* u) {
                                                 if (PyGILState_Check()) {
                                                   py::gil_scoped_release rel;
                                                   delete u;
                                                 } else {
                                                   delete u;
                                                 }
                                               });
           }),
           py::arg("rank"), py::arg("nranks"), py::arg("engine"), py::arg("flows"), py::arg("nlanes") = 2,
           py::arg("tile_bytes") = 1 << 20, py::arg("timeout_ms") = 60000)
      .def_property_readonly("rank", &UkNetComm::rank)
      .def_property_readonly("nranks", &UkNetComm::nranks)
      .def("all_reduce",
           [](UkNetComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype, int op, int algo) {
             py::gil_scoped_release rel;
             u.all_reduce((const void*)in, (void*)out, count, dtype, op, (UkAlgo)algo);
           },
           py::arg("inp"), py::arg("out"), py::arg("count"), py::arg("dtype"), py::arg("op"), py::arg("algo") = 0)
      .def("all_to_all",
           [](UkNetComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype) {
             py::gil_scoped_release rel;
             u.all_to_all((const void*)in, (void*)out, count, dtype);
           })
      .def("all_gather",
           [](UkNetComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype) {
             py::gil_scoped_release rel;
             u.all_gather((const void*)in, (void*)out, count, dtype);
           })
      .def("reduce_scatter",
           [](UkNetComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype, int op) {
             py::gil_scoped_release rel;
             u.reduce_scatter((const void*)in, (void*)out, count, dtype, op);
           })
      .def("broadcast",
           [](UkNetComm& u, uintptr_t in, uintptr_t out, size_t count, int dtype, int root) {
             py::gil_scoped_release rel;
             u.broadcast((const void*)in, (void*)out, count, dtype, root);
           })
      .def("barrier", &UkNetComm::barrier, py::call_guard<py::gil_scoped_release>())
      .def("stats", [](UkNetComm& u) {
        auto s = u.stats();
        py::dict d;
        d["ops"] = s.ops, d["sends"] = s.sends, d["recvs"] = s.recvs, d["bytes_sent"] = s.bytes_sent;
        return d;
      });
}
