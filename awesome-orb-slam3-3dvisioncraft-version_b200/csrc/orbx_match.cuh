// orbx_match.cuh — device-side frame view + scoped stream-ordered allocations used by the matcher /
// optimiser entry points.
#pragma once
#include "orbx_common.cuh"
#include <algorithm>
#include <cstring>
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
  int gridBuilt;           // host-side hint: the grid of this (device-resident) frame is already built
};

// A Frame / KeyFrame kept on the device between calls (include/orbx.h: orbx_frame_upload).
struct orbx_frame {
  orbx_ctx* ctx;
  const void *keyKps, *keyDesc, *keyUright;   // host arrays of the descriptor it was uploaded from
  int n;
  float bounds[4];
  uint8_t* pool;
  FrameDev F;              // device view (grid built)
};

// Temporaries of one host-buffer API call.  Small requests are carved out of the context's staging arena (device
// buffer + page-locked mirror: an upload is one CPU memcpy into the mirror and one truly asynchronous H2D); anything
// that does not fit falls back to the stream-ordered allocator.  The scope holds the context's API mutex; every entry
// point synchronises its stream before returning, so the arena is free again when the next scope starts.
#define ORBX_ARENA_BYTES (24u << 20)
struct DevScope {
  orbx_ctx* ctx;
  cudaStream_t st;
  std::vector<void*> ptrs;
  size_t off = 0;
  bool failed = false;
  DevScope(orbx_ctx* c, cudaStream_t s) : ctx(c), st(s) {
    ctx->apiMutex.lock();
    if (!ctx->arenaDev) {
      if (cudaMalloc(&ctx->arenaDev, ORBX_ARENA_BYTES) == cudaSuccess &&
          cudaHostAlloc(&ctx->arenaHost, ORBX_ARENA_BYTES, cudaHostAllocDefault) == cudaSuccess) {
        ctx->arenaCap = ORBX_ARENA_BYTES;
      } else {
        cudaGetLastError();
        if (ctx->arenaDev) cudaFree(ctx->arenaDev);
        ctx->arenaDev = nullptr;
        ctx->arenaCap = 0;
      }
    }
  }
  ~DevScope() {
    for (void* p : ptrs) cudaFreeAsync(p, st);
    ctx->apiMutex.unlock();
  }
  DevScope(const DevScope&) = delete;
  DevScope& operator=(const DevScope&) = delete;
  // returns the arena offset of a new block, or (size_t)-1 when it does not fit
  size_t carve(size_t bytes) {
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (off + need > ctx->arenaCap) return (size_t)-1;
    const size_t o = off;
    off += need;
    return o;
  }
  template <typename T>
  T* alloc(size_t count) {
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    const size_t o = carve(bytes);
    if (o != (size_t)-1) return (T*)(ctx->arenaDev + o);
    void* p = nullptr;
    if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) {
      failed = true;
      orbx_set_error("orbx: cudaMallocAsync(%zu bytes) failed", bytes);
      return nullptr;
    }
    ptrs.push_back(p);
    return (T*)p;
  }
  // Results: download() enqueues a D2H copy into the page-locked mirror (asynchronous; a pageable destination would make
  // every copy a blocking staged transfer), finish() synchronises the stream once and hands the bytes to the caller.
  struct Pending { void* dst; size_t off, bytes; };
  std::vector<Pending> pending;
  template <typename T>
  void download(T* host_dst, const T* dev_src, size_t count) {
    if (!count || failed) return;
    const size_t bytes = count * sizeof(T);
    const size_t o = carve(bytes);
    cudaError_t e;
    if (o != (size_t)-1) {
      e = cudaMemcpyAsync(ctx->arenaHost + o, dev_src, bytes, cudaMemcpyDeviceToHost, st);
      pending.push_back(Pending{(void*)host_dst, o, bytes});
    } else {
      e = cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, st);
    }
    if (e != cudaSuccess) { failed = true; orbx_set_error("orbx: D2H copy failed: %s", cudaGetErrorString(e)); }
  }
  int finish() {
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess || failed) {
      if (e != cudaSuccess) orbx_set_error("orbx: %s", cudaGetErrorString(e));
      return ORBX_ECUDA;
    }
    for (const Pending& p : pending) memcpy(p.dst, ctx->arenaHost + p.off, p.bytes);
    pending.clear();
    return ORBX_OK;
  }
  template <typename T>
  T* upload(const T* host, size_t count) {
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    const size_t o = carve(bytes);
    if (o != (size_t)-1) {
      T* d = (T*)(ctx->arenaDev + o);
      if (count) {
        memcpy(ctx->arenaHost + o, host, count * sizeof(T));
        if (cudaMemcpyAsync(d, ctx->arenaHost + o, count * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess) {
          failed = true;
          orbx_set_error("orbx: H2D copy failed");
        }
      }
      return d;
    }
    T* d = alloc<T>(count);
    if (d && count && cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess) {
      failed = true;
      orbx_set_error("orbx: H2D copy failed");
    }
    return d;
  }
};

// Layout of a persistent device pool built on the host first: add() copies bytes into the host image and returns their
// offset, reserve() leaves room for device-side scratch / outputs; the owner then does ONE cudaMalloc(size()) + ONE
// cudaMemcpy and turns offsets into pointers.  Used by the prepared many-problem plans (orbx_tri_batch, orbx_lba_batch).
struct PoolBuilder {
  std::vector<uint8_t> h;
  size_t reserve(size_t bytes) {
    const size_t o = (h.size() + 255) & ~(size_t)255;
    h.resize(o + std::max<size_t>(bytes, 1), 0);
    return o;
  }
  size_t add(const void* src, size_t bytes) {
    const size_t o = reserve(bytes);
    if (bytes) memcpy(h.data() + o, src, bytes);
    return o;
  }
  template <typename T> size_t addv(const std::vector<T>& v) { return add(v.data(), sizeof(T) * v.size()); }
  size_t size() const { return (h.size() + 255) & ~(size_t)255; }
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
  // row index of the right keypoints (stereo_rows_kernel): sortIdx = right keypoint ids bucketed by floor(y), rowStart[r] =
  // first entry of row r (rowStart[rows] = nR); rowStart[-1 + 0] < 0 marks "not built" (image taller than the kernel's table)
  uint16_t* sortIdx;   // [nR] or null: scan every right keypoint
  int* rowStart;       // [ORBX_STEREO_MAX_ROWS + 1]
};
#define ORBX_STEREO_MAX_ROWS 2304

// batched launchers (S frames; arg/frame arrays are device pointers)
int orbx_launch_stereo_batch(orbx_ctx* ctx, cudaStream_t st, const StereoArgs* dArgs, int S, int maxL);
int orbx_launch_sbp_frame_batch(orbx_ctx* ctx, cudaStream_t st, const FrameDev* dF, const SbpFrameArgs* dA, int S, int maxQ,
                                int maxN);
int orbx_launch_sbp_map_batch(orbx_ctx* ctx, cudaStream_t st, const FrameDev* dF, const SbpMapArgs* dA, int S, int maxQ,
                              int maxN);
