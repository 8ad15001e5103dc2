// orbx_match.cuh — device-side frame view + scoped stream-ordered allocations used by the matcher /
// optimiser entry points.
#pragma once
#include "orbx_common.cuh"
#include <algorithm>
#include <vector>

#define ORBX_NCELLS (ORBX_GRID_COLS * ORBX_GRID_ROWS)

// One Frame / KeyFrame as the kernels see it (all pointers are device pointers).
struct FrameDev {
  int n;                   // keypoint count (host-known) ...
  const int* nDev;         // ... or, when non-null, a device counter that holds it (batched pipelines)
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

// ---- per-frame argument blocks of the matcher kernels (arrays of these drive batched launches) ----
struct SbpMapArgs {
  int nq;
  const int* nqDev;     // when non-null: device-resident query count (batched pipelines)
  const float *projX, *projY, *projXR, *viewCos;
  const int* level;
  const uint8_t* mpDesc;
  const uint8_t* flags;
  float th, nnratio;
  const float* scaleFactors;
  // phase A -> B
  int* candOfs;      // [nq]
  int* candCnt;      // [nq]
  uint32_t* cand;    // [candCap]
  int candCap;
  int* total;        // running allocation counter
  int* err;
  // outputs
  const uint8_t* kpBlocked;
  int* bestIdx;
  int* nmatches;
};

struct SbpFrameArgs {
  int nq;
  const int* nqDev;
  const float* TcDev;   // when non-null: device pose [12+] overriding Tc (batched pipelines)
  const uint8_t* flags;
  const float* xw;
  const int* octave;
  const float* angle;
  const uint8_t* mpDesc;
  float Tc[12];
  float fx, fy, cx, cy, bf;
  float th;
  int mode;          // 0 = +-1 octave, 1 = forward, 2 = backward
  int checkOri;
  const float* scaleFactors;
  int* candOfs;
  int* candCnt;
  uint32_t* cand;    // idx | dist<<16
  int candCap;
  int* total;
  int* err;
  const uint8_t* curBlocked;
  int* matchIdx;
  uint8_t* kept;
  int* curMatch;
  int* nmatches;
};

struct StereoArgs {
  int nL, nR;
  const int *nLDev, *nRDev;   // when non-null: device-resident keypoint counts
  const orbx_keypoint *kpL, *kpR;
  const uint8_t *descL, *descR;
  float bf, b;
  int nlevels;
  float scale[ORBX_MAX_LEVELS], invScale[ORBX_MAX_LEVELS];
  const uint8_t* pyrL[ORBX_MAX_LEVELS];
  const uint8_t* pyrR[ORBX_MAX_LEVELS];
  int lw[ORBX_MAX_LEVELS], lh[ORBX_MAX_LEVELS], pitchL[ORBX_MAX_LEVELS], pitchR[ORBX_MAX_LEVELS];
  float* uright;
  float* depth;
  int* sad;      // [nL] best SAD of accepted matches, -1 otherwise
};

// batched launchers (S frames; arg/frame arrays are device pointers)
int orbx_launch_stereo_batch(orbx_ctx* ctx, cudaStream_t st, const StereoArgs* dArgs, int S, int maxL);
int orbx_launch_sbp_frame_batch(orbx_ctx* ctx, cudaStream_t st, const FrameDev* dF, const SbpFrameArgs* dA, int S, int maxQ,
                                int maxN);
int orbx_launch_sbp_map_batch(orbx_ctx* ctx, cudaStream_t st, const FrameDev* dF, const SbpMapArgs* dA, int S, int maxQ,
                              int maxN);
