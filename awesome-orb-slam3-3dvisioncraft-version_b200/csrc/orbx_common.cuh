// orbx_common.cuh — shared host/device plumbing for liborbx.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <atomic>
#include <mutex>
#include <vector>
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
  // Call-scoped staging arena of the host-buffer entry points (DevScope): one device buffer + its page-locked mirror,
  // bump-allocated, so the ~20 small uploads / temporaries of a matcher or optimiser call cost one pinned memcpy each
  // instead of a cudaMallocAsync + pageable copy.  apiMutex serialises those calls per context (they serialise on the
  // context's stream anyway).
  std::mutex apiMutex;
  uint8_t* arenaDev = nullptr;
  uint8_t* arenaHost = nullptr;
  size_t arenaCap = 0;
  // Device-resident frames (orbx_frame_upload): looked up by the matcher entry points under apiMutex, keyed by the host
  // arrays of the descriptor they were uploaded from.
  std::vector<struct orbx_frame*> residentFrames;
};

// Device-resident arguments of the visual-inertial pose optimisers when the tracker drives them (orbx_optimize.cu:
// orbx_launch_pose_inertial_slices).  Every pointer is device memory; per-stream arrays carry a leading [S] dimension.
struct OrbxInertialSlices {
  int mode;                 // 1: PoseInertialOptimizationLastKeyFrame, 2: PoseInertialOptimizationLastFrame
  int S, recInit;
  float fx, fy, cx, cy, bf;
  const int *estart, *ecount;               // edge slices
  const float *exw, *eobs, *eisg;
  const uint8_t* eclose;
  uint8_t* eoutlier;
  double* err;                              // [3 * edges] scratch
  const float* T1;                          // [S][16] frame pose before the optimisation (pFrame->mTcw)
  const float *Tcb, *Tbc;                   // [16]
  const float* vel;                         // [S][3] pFrame->mVw
  const float* bias;                        // [S][6] pFrame->mImuBias, gyro xyz then acc xyz
  const double *ref, *preint, *preintJac, *preintBias, *infoI, *infoG, *infoA, *priorState, *priorH;
  double *stateOut, *H15;                   // [S][21], [S][225]
  int *nRet, *iters;                        // [S], [S][4]
  float* T2;                                // [S][16] pose after Frame::SetImuPoseVelocity
};

#define ORBX_LAUNCH(ctx) ((ctx)->launches.fetch_add(1, std::memory_order_relaxed))

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
