// Placeholder binders; replaced as subsystems land.
#include <pybind11/pybind11.h>
namespace py = pybind11;
#ifndef UB_HAVE_EP
void bind_ep(py::module_&) {}
#endif
#ifndef UB_HAVE_P2P
void bind_p2p(py::module_&) {}
#endif
#ifndef UB_HAVE_UTIL
void bind_util(py::module_&) {}
#endif
#ifndef UB_HAVE_UK
void bind_uk(py::module_&) {}
#endif
