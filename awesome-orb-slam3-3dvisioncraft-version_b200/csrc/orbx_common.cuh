// orbx_common.cuh — shared host/device plumbing for liborbx.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <atomic>
#include "../../include/orbx.h"

void orbx_set_error(const char* fmt, ...);

#define ORBX_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      orbx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));   \
      return ORBX_ECUDA;                                                                      \
    }                                                                                         \
  } while (0)

struct orbx_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::atomic<uint64_t> launches{0};
};

#define ORBX_LAUNCH(ctx) ((ctx)->launches.fetch_add(1, std::memory_order_relaxed))

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
