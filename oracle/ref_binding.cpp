// pybind11 module exposing the reference's four hash-grid entry points, declared in
// /root/reference/wisp/csrc/ops/hashgrid_interpolate.h:18-50 (the same attribute set the
// reference registers as wisp._C.ops in wisp/csrc/bindings.cpp:23-27). Built by
// oracle/build_ref.py into oracle/_ref/; the reference sources are compiled in place and
// never copied. Test infrastructure only (REF-GPU oracle / baseline).
#include <torch/extension.h>
#include "hashgrid_interpolate.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("hashgrid_interpolate_cuda", &wisp::hashgrid_interpolate_cuda);
    m.def("hashgrid_interpolate_backward_cuda", &wisp::hashgrid_interpolate_backward_cuda);
    m.def("hashgrid_interpolate2d_cuda", &wisp::hashgrid_interpolate2d_cuda);
    m.def("hashgrid_interpolate2d_backward_cuda", &wisp::hashgrid_interpolate2d_backward_cuda);
}
