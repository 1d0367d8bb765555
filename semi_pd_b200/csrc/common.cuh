// Shared host-side helpers for libspd.so: error reporting, argument checks, dtype dispatch.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/spd.h"

namespace spd {

// thread-local last-error text (spd_last_error_string)
char* last_error_buf();
inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define SPD_REQUIRE(cond, code, ...)                 \
  do {                                               \
    if (!(cond)) return ::spd::set_error(code, __VA_ARGS__); \
  } while (0)

// check the launch that was just issued (no sync)
#define SPD_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess)                                                           \
      return ::spd::set_error(static_cast<int>(e__), "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define SPD_CUDA(call)                                                                \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return ::spd::set_error(static_cast<int>(e__), "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

// dtype dispatch: runs `...` with `T` bound to __nv_bfloat16 or __half
#define SPD_DISPATCH_DTYPE(dtype, ...)                                  \
  do {                                                                  \
    if ((dtype) == SPD_BF16) {                                          \
      using T = __nv_bfloat16;                                          \
      __VA_ARGS__                                                       \
    } else if ((dtype) == SPD_F16) {                                    \
      using T = __half;                                                 \
      __VA_ARGS__                                                       \
    } else {                                                            \
      return ::spd::set_error(SPD_E_BADARG, "bad dtype %d", (dtype));   \
    }                                                                   \
  } while (0)

inline cudaStream_t as_stream(spd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int device_sm_count_cached();
// programmatic dependent launch for the per-layer kernels (SPD_PDL=0 in the environment disables it)
bool pdl_enabled();

// launch helper: optional programmatic-dependent-launch attribute.  Every kernel launched through
// it executes griddepcontrol.wait before its first global-memory access, so a chain of such
// launches is ordered exactly like plain stream order.
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          cudaStream_t stream, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace spd
