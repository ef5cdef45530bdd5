// Python bindings of the EP runtime (role of the reference's NB_MODULE in ep/src/uccl_ep.cc:1641-2410).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "ep_buffer.h"

namespace py = pybind11;
using namespace ub;

void bind_ep(py::module_& m) {
  m.attr("EP_X_BF16") = (int)EP_X_BF16;
  m.attr("EP_X_FP8_SCALED") = (int)EP_X_FP8_SCALED;
  m.attr("EP_X_FUSED_FP8") = (int)EP_X_FUSED_FP8;
  m.attr("EP_MAX_TOPK") = kEpMaxTopk;
  py::class_<EpDispatchOut>(m, "EpDispatchOut")
      .def_readonly("recv_x", &EpDispatchOut::recv_x)
      .def_readonly("recv_scales", &EpDispatchOut::recv_scales)
      .def_readonly("recv_topk_idx", &EpDispatchOut::recv_topk_idx)
      .def_readonly("recv_topk_w", &EpDispatchOut::recv_topk_w)
      .def_readonly("recv_src_idx", &EpDispatchOut::recv_src_idx)
      .def_readonly("slot", &EpDispatchOut::slot)
      .def_readonly("capacity", &EpDispatchOut::capacity);
  py::class_<EpBuffer, std::shared_ptr<EpBuffer>>(m, "EpBuffer")
      .def(py::init<std::shared_ptr<Comm>, size_t, int>(), py::arg("comm"), py::arg("num_nvl_bytes"),
           py::arg("num_slots") = 2)
      .def_property_readonly("rank", &EpBuffer::rank)
      .def_property_readonly("nranks", &EpBuffer::nranks)
      .def_property_readonly("arena_bytes", &EpBuffer::arena_bytes)
      .def_property_readonly("num_slots", &EpBuffer::num_slots)
      .def_property_readonly("launches", &EpBuffer::launches)
      .def_property_readonly("dev_counts_ptr", &EpBuffer::dev_counts_ptr)
      .def("capacity_for", &EpBuffer::capacity_for)
      .def("combine_capacity_for", &EpBuffer::combine_capacity_for)
      .def("layout",
           [](EpBuffer& b, uintptr_t topk_idx, int T, int K, int E, uintptr_t tpr, uintptr_t tpe, uintptr_t itir,
              uintptr_t pos, uintptr_t st) { b.layout(topk_idx, T, K, E, tpr, tpe, itir, pos, (cudaStream_t)st); })
      .def("dispatch",
           [](EpBuffer& b, uintptr_t x, uintptr_t xs, uintptr_t ti, uintptr_t tw, uintptr_t pos, uintptr_t ss,
              uintptr_t tpr, uintptr_t tpe, int T, int H, int K, int E, int mode, bool cached, int reuse_slot,
              uintptr_t rank_prefix, int expert_alignment, int num_worst_tokens, bool round_scale, int num_sms,
              uintptr_t st) {
             return b.dispatch(x, xs, ti, tw, pos, ss, tpr, tpe, T, H, K, E, mode, cached, reuse_slot, rank_prefix,
                               expert_alignment, num_worst_tokens, round_scale, num_sms, (cudaStream_t)st);
           })
      .def("wait_counts",
           [](EpBuffer& b, int E_local, double timeout_s) {
             std::vector<int> pe;
             int total;
             {
               py::gil_scoped_release rel;
               total = b.wait_counts(E_local, &pe, timeout_s);
             }
             return py::make_tuple(total, pe);
           },
           py::arg("E_local"), py::arg("timeout_s") = 0.0)
      .def("ll_init", &EpBuffer::ll_init)
      .def_static("ll_size_hint", &EpBuffer::ll_size_hint)
      .def("ll_dispatch",
           [](EpBuffer& b, uintptr_t x, uintptr_t ti, int T, int H, int K, int E, int M, bool use_fp8, bool round_scale,
              uintptr_t recv_count, uintptr_t layout_range, uintptr_t send_pos, int num_sms, uintptr_t st) {
             auto o = b.ll_dispatch(x, ti, T, H, K, E, M, use_fp8, round_scale, recv_count, layout_range, send_pos,
                                    num_sms, (cudaStream_t)st);
             return py::make_tuple(o.recv_x, o.recv_scales, o.recv_src_info, o.combine_x, o.buffer_idx);
           })
      .def("ll_combine_buffer", &EpBuffer::ll_combine_buffer)
      .def("ll_combine",
           [](EpBuffer& b, uintptr_t x, int idx, uintptr_t tw, uintptr_t sp, uintptr_t out, int T, int H, int K, int E,
              int M, int num_sms, uintptr_t st) {
             b.ll_combine(x, idx, tw, sp, out, T, H, K, E, M, num_sms, (cudaStream_t)st);
           })
      .def("combine_input_ptr", &EpBuffer::combine_input_ptr)
      .def("combine", [](EpBuffer& b, uintptr_t x, int num_recv, uintptr_t tw, uintptr_t ss, uintptr_t b0,
                         uintptr_t b1, uintptr_t out, uintptr_t otw, int T, int H, int K, int num_sms, uintptr_t st) {
        b.combine(x, num_recv, tw, ss, b0, b1, out, otw, T, H, K, num_sms, (cudaStream_t)st);
      });
}
