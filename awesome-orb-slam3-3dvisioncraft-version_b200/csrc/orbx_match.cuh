// orbx_match.cuh — device-side frame view + scoped stream-ordered allocations used by the matcher /
// optimiser entry points.
#pragma once
#include "orbx_common.cuh"
#include <vector>

#define ORBX_NCELLS (ORBX_GRID_COLS * ORBX_GRID_ROWS)

// One Frame / KeyFrame as the kernels see it (all pointers are device pointers).
struct FrameDev {
  int n;
  const orbx_keypoint* kps;
  const uint8_t* desc;
  const float* uright;     // may be null
  float minX, minY, maxX, maxY, wInv, hInv;
  int* cellStart;          // [ORBX_NCELLS + 1]  CSR of the 64x48 grid, cell id = ix * 48 + iy
  int* cellIdx;            // [n] keypoint indices, ascending inside a cell
};

// Stream-ordered temporaries of one API call (cudaMallocAsync pool; freed on scope exit).
struct DevScope {
  cudaStream_t st;
  std::vector<void*> ptrs;
  bool failed = false;
  explicit DevScope(cudaStream_t s) : st(s) {}
  ~DevScope() {
    for (void* p : ptrs) cudaFreeAsync(p, st);
  }
  template <typename T>
  T* alloc(size_t count) {
    void* p = nullptr;
    if (cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), st) != cudaSuccess) {
      failed = true;
      orbx_set_error("orbx: cudaMallocAsync(%zu bytes) failed", count * sizeof(T));
      return nullptr;
    }
    ptrs.push_back(p);
    return (T*)p;
  }
  template <typename T>
  T* upload(const T* host, size_t count) {
    T* d = alloc<T>(count);
    if (d && count && cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess) {
      failed = true;
      orbx_set_error("orbx: H2D copy failed");
    }
    return d;
  }
};

// Upload one orbx_frame_desc and allocate its grid; fills `out` (host copy of the device view).
int orbx_upload_frame(DevScope& S, const orbx_frame_desc* f, FrameDev* out);
// grid build for nFrames frames (d_frames = device array); one CTA per frame
int orbx_launch_grid_build(orbx_ctx* ctx, cudaStream_t st, const FrameDev* d_frames, int nFrames);
