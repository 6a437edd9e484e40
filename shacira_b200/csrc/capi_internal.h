// capi_internal.h -- error reporting, launch counting and level metadata shared by the C-ABI
// translation units (capi.cu, tiled_capi.cu). Internal; the public surface is include/shacira_b200.h.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace shacira {

char* last_error_buffer();           // thread-local, 512 bytes
std::atomic<int64_t>& launch_counter();

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver         \
                            ? SHACIRA_ERR_NO_DEVICE : SHACIRA_ERR_CUDA,                        \
                        "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define LAUNCHED()                            \
    do {                                      \
        launch_counter().fetch_add(1);        \
        CUDA_OK(cudaGetLastError());          \
    } while (0)

// Level metadata + the checks the reference does not make (SURVEY 8b "error convention").
int build_levels(int32_t dim, const int32_t* first_idx, const int32_t* resolutions, int32_t num_lods,
                 int32_t bitwidth, LevelParams& lp);
int check_points(const float* coords, int64_t n);

}  // namespace shacira
