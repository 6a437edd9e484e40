// Pre-included (-include) when compiling the reference's UNMODIFIED .cu files where they
// lie under /root/reference/wisp/csrc/ops. torch >= 2.x removed the
// AT_DISPATCH_*(tensor.type(), ...) overload the reference still uses
// (hashgrid_interpolate_cuda.cu:125,290; hashgrid_interpolate2d_cuda.cu:115,251); this adds
// it back so no reference source has to be copied or patched. Test infrastructure only.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
