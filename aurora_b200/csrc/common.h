// Host-side helpers shared by every translation unit of libaurora_b200.so.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/aurora_b200.h"

namespace ab {

// Kernel-launch counter exposed through ab_launch_count().
extern std::atomic<unsigned long long> g_launches;
#define AB_COUNT_LAUNCH(n) ::ab::g_launches.fetch_add((n), std::memory_order_relaxed)

// Thread-local last-error string returned by ab_last_error().
void set_error(const char* fmt, ...);

// Number of SMs of the current device (cached per device).
int sm_count();

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();

// 2D bf16 / fp16 tensor map: inner dim = `cols` contiguous elements, outer dim = `rows` with `ld` elements
// between rows; box = {box_cols, box_rows}; 128-byte swizzle; out-of-bounds reads return zero.
int make_tmap_16bit_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols, bool fp16);

#define AB_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::ab::set_error(__VA_ARGS__);        \
      return AB_ERR_INVALID_ARGUMENT;      \
    }                                      \
  } while (0)

#define AB_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      ::ab::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));       \
      return AB_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace ab
