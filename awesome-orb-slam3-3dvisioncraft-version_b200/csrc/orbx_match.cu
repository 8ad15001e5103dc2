// orbx_match.cu — sm_100a descriptor matchers behind include/orbx.h.
//
// Replaces (reference paths): Frame::AssignFeaturesToGrid/GetFeaturesInArea src/Frame.cc:444-478,755-850;
// Frame::ComputeStereoMatches src/Frame.cc:955-1133; ORBmatcher::SearchByProjection x2
// src/ORBmatcher.cc:59-255,2244-2509; SearchForTriangulation :1138-1428; ComputeThreeMaxima :2654-2695;
// DescriptorDistance :2700-2716; Pinhole::epipolarConstrain src/CameraModels/Pinhole.cpp:155-177.
//
// Structure shared by both projection searches: the expensive part (grid walk + 256-bit Hamming via
// __popc over 8 words) runs one warp per query, fully parallel, and records each query's candidates in
// the reference's visiting order; the reference's *ordered greedy* semantics (a keypoint taken by an
// earlier MapPoint with observations is skipped by later ones) is then replayed exactly by a single warp
// per frame walking the queries in order with warp-wide (dist,position) min-reductions.
#include <algorithm>
#include "orbx_match.cuh"

#define TH_HIGH 100
#define TH_LOW 50
#define HISTO_LENGTH 30

// ------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int hamming256(const uint4* __restrict__ a, const uint4* __restrict__ b) {
  const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// C round(): half away from zero
__device__ __forceinline__ int round_haz(float v) { return (int)roundf(v); }

struct CellRange { int x0, x1, y0, y1; bool empty; };

// frame f of a batch, with its keypoint count resolved
__device__ __forceinline__ FrameDev load_frame(const FrameDev* frames, int f) {
  FrameDev F = frames[f];
  if (F.nDev) F.n = *F.nDev;
  return F;
}

// GetFeaturesInArea cell window (src/Frame.cc:779-802)
__device__ __forceinline__ CellRange cell_range(const FrameDev& F, float x, float y, float r) {
  CellRange c;
  c.empty = true;
  c.x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, F.minX), r), F.wInv)));
  if (c.x0 >= ORBX_GRID_COLS) return c;
  c.x1 = min(ORBX_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, F.minX), r), F.wInv)));
  if (c.x1 < 0) return c;
  c.y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, F.minY), r), F.hInv)));
  if (c.y0 >= ORBX_GRID_ROWS) return c;
  c.y1 = min(ORBX_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, F.minY), r), F.hInv)));
  if (c.y1 < 0) return c;
  c.empty = false;
  return c;
}

__device__ __forceinline__ bool in_window(const orbx_keypoint& kp, float x, float y, float r, int minLevel, int maxLevel) {
  const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
  if (checkLevels) {
    if (kp.octave < minLevel) return false;
    if (maxLevel >= 0 && kp.octave > maxLevel) return false;
  }
  return fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r;
}

// Warp-cooperative walk of the candidates of one query in the reference's order.  Calls
// f(position, keypointIndex) for every keypoint GetFeaturesInArea would return; `position` is the
// index in the returned vector.  All 32 lanes must call; f runs on the lane that owns the candidate.
// Returns the total number of candidates.
template <typename Fn>
__device__ __forceinline__ int walk_candidates(const FrameDev& F, float x, float y, float r, int minLevel, int maxLevel,
                                               Fn f) {
  const int lane = threadIdx.x & 31;
  const CellRange c = cell_range(F, x, y, r);
  if (c.empty) return 0;
  int pos = 0;
  for (int ix = c.x0; ix <= c.x1; ++ix) {
    // cells (ix, y0..y1) are contiguous in the CSR (cell id = ix*48 + iy): one linear span per column
    const int beg = F.cellStart[ix * ORBX_GRID_ROWS + c.y0], end = F.cellStart[ix * ORBX_GRID_ROWS + c.y1 + 1];
    for (int base = beg; base < end; base += 32) {
      const int i = base + lane;
      int idx = -1;
      bool ok = false;
      if (i < end) {
        idx = F.cellIdx[i];
        ok = in_window(F.kps[idx], x, y, r, minLevel, maxLevel);
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) f(pos + __popc(m & ((1u << lane) - 1)), idx);
      pos += __popc(m);
    }
  }
  return pos;
}

// walk_candidates for the two-phase score kernels: the gated candidate of every (column, chunk) iteration is remembered per
// lane (idx, or -1), so that the fill phase -- which needs the query's total first, to reserve its slice of the candidate
// list -- replays the iterations from this cache instead of walking the grid, the keypoints and the uRight gate a second
// time.  nIt = -1 when the walk has more than SBP_CACHE iterations (the caller then walks again).
#define SBP_CACHE 24
template <typename Gate>
__device__ __forceinline__ int walk_cached(const FrameDev& F, float x, float y, float r, int minLevel, int maxLevel, Gate gate,
                                           int (&cache)[SBP_CACHE], int& nIt) {
  const int lane = threadIdx.x & 31;
  const CellRange c = cell_range(F, x, y, r);
  nIt = 0;
  if (c.empty) return 0;
  int cnt = 0, it = 0;
  for (int ix = c.x0; ix <= c.x1; ++ix) {
    const int beg = F.cellStart[ix * ORBX_GRID_ROWS + c.y0], end = F.cellStart[ix * ORBX_GRID_ROWS + c.y1 + 1];
    for (int base = beg; base < end; base += 32, ++it) {
      const int i = base + lane;
      int idx = -1;
      if (i < end) {
        idx = F.cellIdx[i];
        if (!(in_window(F.kps[idx], x, y, r, minLevel, maxLevel) && gate(idx))) idx = -1;
      }
      if (it < SBP_CACHE) cache[it] = idx;
      cnt += idx >= 0;
    }
  }
  nIt = it <= SBP_CACHE ? it : -1;
  return cnt;
}

// ------------------------------------------------------------------------------------
// K7 grid build: one CTA per frame.  Counting sort by cell, ascending keypoint index in a cell.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_build_kernel(const FrameDev* __restrict__ frames) {
  const FrameDev F = load_frame(frames, blockIdx.x);
  __shared__ int s_cnt[ORBX_NCELLS];
  __shared__ int s_warp[9];
  const int tid = threadIdx.x;
  for (int i = tid; i < ORBX_NCELLS; i += 256) s_cnt[i] = 0;
  __syncthreads();
  for (int i = tid; i < F.n; i += 256) {
    const orbx_keypoint kp = F.kps[i];
    const int px = round_haz(__fmul_rn(__fsub_rn(kp.x, F.minX), F.wInv));   // PosInGrid, src/Frame.cc:852-862
    const int py = round_haz(__fmul_rn(__fsub_rn(kp.y, F.minY), F.hInv));
    if (px < 0 || px >= ORBX_GRID_COLS || py < 0 || py >= ORBX_GRID_ROWS) continue;
    atomicAdd(&s_cnt[px * ORBX_GRID_ROWS + py], 1);
  }
  __syncthreads();
  // exclusive scan of 3072 counts: 12 per thread
  const int per = ORBX_NCELLS / 256;
  int local[ORBX_NCELLS / 256];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < per; ++k) { local[k] = s_cnt[tid * per + k]; sum += local[k]; }
  int incl = sum;
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (tid == 0) { int run = 0; for (int w = 0; w < 8; ++w) { int v = s_warp[w]; s_warp[w] = run; run += v; } s_warp[8] = run; }
  __syncthreads();
  int run = s_warp[wid] + incl - sum;
#pragma unroll
  for (int k = 0; k < per; ++k) { F.cellStart[tid * per + k] = run; s_cnt[tid * per + k] = run; run += local[k]; }
  if (tid == 255) F.cellStart[ORBX_NCELLS] = s_warp[8];
  __syncthreads();
  // unordered fill, then sort each cell's few entries ascending (= insertion order of the reference)
  for (int i = tid; i < F.n; i += 256) {
    const orbx_keypoint kp = F.kps[i];
    const int px = round_haz(__fmul_rn(__fsub_rn(kp.x, F.minX), F.wInv));
    const int py = round_haz(__fmul_rn(__fsub_rn(kp.y, F.minY), F.hInv));
    if (px < 0 || px >= ORBX_GRID_COLS || py < 0 || py >= ORBX_GRID_ROWS) continue;
    F.cellIdx[atomicAdd(&s_cnt[px * ORBX_GRID_ROWS + py], 1)] = i;
  }
  __syncthreads();
  for (int c = tid; c < ORBX_NCELLS; c += 256) {
    const int b = F.cellStart[c], e = s_cnt[c];
    for (int i = b + 1; i < e; ++i) {
      const int v = F.cellIdx[i];
      int j = i - 1;
      while (j >= b && F.cellIdx[j] > v) { F.cellIdx[j + 1] = F.cellIdx[j]; --j; }
      F.cellIdx[j + 1] = v;
    }
  }
}

int orbx_launch_grid_build(orbx_ctx* ctx, cudaStream_t st, const FrameDev* d_frames, int nFrames) {
  grid_build_kernel<<<nFrames, 256, 0, st>>>(d_frames);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

int orbx_upload_frame(DevScope& S, const orbx_frame_desc* f, FrameDev* out) {
  if (!f || f->n < 0 || (f->n > 0 && (!f->kps || !f->desc)) || !(f->max_x > f->min_x) || !(f->max_y > f->min_y)) {
    orbx_set_error("orbx: invalid frame descriptor");
    return ORBX_EINVAL;
  }
  // a frame that was made device-resident with orbx_frame_upload: nothing to copy, grid already built
  for (orbx_frame* R : S.ctx->residentFrames)
    if (R->keyKps == (const void*)f->kps && R->keyDesc == (const void*)f->desc && R->keyUright == (const void*)f->uright && R->n == f->n &&
        R->bounds[0] == f->min_x && R->bounds[1] == f->min_y && R->bounds[2] == f->max_x && R->bounds[3] == f->max_y) {
      *out = R->F;
      return ORBX_OK;
    }
  out->n = f->n;
  out->nDev = nullptr;
  out->kps = S.upload(f->kps, f->n);
  out->desc = S.upload(f->desc, (size_t)f->n * 32);
  out->uright = f->uright ? S.upload(f->uright, f->n) : nullptr;
  out->minX = f->min_x;
  out->minY = f->min_y;
  out->maxX = f->max_x;
  out->maxY = f->max_y;
  out->wInv = (float)ORBX_GRID_COLS / (f->max_x - f->min_x);
  out->hInv = (float)ORBX_GRID_ROWS / (f->max_y - f->min_y);
  out->cellStart = S.alloc<int>(ORBX_NCELLS + 1);
  out->cellIdx = S.alloc<int>(f->n);
  out->gridBuilt = 0;
  return S.failed ? ORBX_ECUDA : ORBX_OK;
}

// ------------------------------------------------------------------------------------
// features_in_area (test primitive): warp per query
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) features_in_area_kernel(const FrameDev* frames, int nq, const float* x, const float* y,
                                                               const float* r, const int* minL, const int* maxL,
                                                               int* outIdx, int cap, int* outN) {
  const FrameDev F = load_frame(frames, 0);
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  int* dst = outIdx + (size_t)q * cap;
  const int n = walk_candidates(F, x[q], y[q], r[q], minL[q], maxL[q], [&](int pos, int idx) {
    if (pos < cap) dst[pos] = idx;
  });
  if ((threadIdx.x & 31) == 0) outN[q] = n;
}

// ------------------------------------------------------------------------------------
// K8a  SearchByProjection(Frame, MapPoints): phase A (parallel candidate scoring)
//   candidate record: idx | dist<<16 | (level & 0xf) << 25
// ------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) sbp_map_score_kernel(const FrameDev* frames, const SbpMapArgs* args) {
  const SbpMapArgs A = args[blockIdx.y];
  const FrameDev F = load_frame(frames, blockIdx.y);
  const int nq = A.nqDev ? *A.nqDev : A.nq;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= nq) return;
  if (lane == 0) { A.candCnt[q] = 0; A.candOfs[q] = 0; }
  if (!(A.flags[q] & 1)) return;
  const int lvl = A.level[q];
  float r = ((double)A.viewCos[q] > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos, src/ORBmatcher.cc:260-266
  if (A.th != 1.0f) r = __fmul_rn(r, A.th);
  const float rs = __fmul_rn(r, A.scaleFactors[lvl]);
  const float x = A.projX[q], y = A.projY[q];
  // pass 1: count (static gates only; the "already assigned" gate is dynamic and applied in phase B)
  const float xr = A.projXR ? A.projXR[q] : 0.f;
  auto gate = [&](int idx) {
    if (F.uright) {
      const float ur = F.uright[idx];
      if (ur > 0 && fabsf(__fsub_rn(xr, ur)) > rs) return false;
    }
    return true;
  };
  int cache[SBP_CACHE], nIt;
  int cnt = walk_cached(F, x, y, rs, lvl - 1, lvl, gate, cache, nIt);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (cnt == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(A.total, cnt);
  base = __shfl_sync(0xffffffffu, base, 0);
  if (base + cnt > A.candCap) {
    if (lane == 0) atomicExch(A.err, 1);
    return;
  }
  // pass 2: fill in visiting order.  Position among *gated* candidates = rank of its walk position.
  const uint4* dq = reinterpret_cast<const uint4*>(A.mpDesc + 32 * (size_t)q);
  const CellRange c = cell_range(F, x, y, rs);
  int pos = 0;
  for (int it = 0; it < nIt; ++it) {                     // replay of the cached walk (nIt = -1: walk again below)
    const int idx = cache[it];
    const bool ok = idx >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int d = hamming256(dq, reinterpret_cast<const uint4*>(F.desc + 32 * (size_t)idx));
      A.cand[base + pos + __popc(m & ((1u << lane) - 1))] =
          (uint32_t)idx | ((uint32_t)d << 16) | ((uint32_t)(F.kps[idx].octave & 0xf) << 25);
    }
    pos += __popc(m);
  }
  for (int ix = c.x0; nIt < 0 && ix <= c.x1; ++ix) {
    const int beg = F.cellStart[ix * ORBX_GRID_ROWS + c.y0], end = F.cellStart[ix * ORBX_GRID_ROWS + c.y1 + 1];
    for (int b0 = beg; b0 < end; b0 += 32) {
      const int i = b0 + lane;
      int idx = -1;
      bool ok = false;
      if (i < end) {
        idx = F.cellIdx[i];
        ok = in_window(F.kps[idx], x, y, rs, lvl - 1, lvl) && gate(idx);
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int d = hamming256(dq, reinterpret_cast<const uint4*>(F.desc + 32 * (size_t)idx));
        A.cand[base + pos + __popc(m & ((1u << lane) - 1))] =
            (uint32_t)idx | ((uint32_t)d << 16) | ((uint32_t)(F.kps[idx].octave & 0xf) << 25);
      }
      pos += __popc(m);
    }
  }
  if (lane == 0) { A.candCnt[q] = cnt; A.candOfs[q] = base; }
}

// phase B: one warp per frame replays the queries in order
__global__ void __launch_bounds__(32) sbp_map_resolve_kernel(const FrameDev* frames, const SbpMapArgs* args) {
  extern __shared__ uint8_t s_blocked[];
  const SbpMapArgs A = args[blockIdx.x];
  const FrameDev F = load_frame(frames, blockIdx.x);
  const int nq = A.nqDev ? *A.nqDev : A.nq;
  const int lane = threadIdx.x;
  for (int i = lane; i < F.n; i += 32) s_blocked[i] = A.kpBlocked ? A.kpBlocked[i] : 0;
  __syncwarp();
  // The replay is a serial chain over the queries; what made it slow was three dependent global loads per query
  // (count -> offset -> candidates).  Counts/offsets/flags of 32 queries are fetched with one coalesced load each and
  // the first 32 candidates of 8 queries are prefetched together, so the chain itself only touches registers and
  // shared memory.
  const unsigned FULL = 0xffffffffu;
  int n = 0;
  for (int q0 = 0; q0 < nq; q0 += 32) {
    const int qq = q0 + lane;
    const int myCnt = qq < nq ? A.candCnt[qq] : 0;
    const int myOfs = qq < nq ? A.candOfs[qq] : 0;
    const int myFlag = qq < nq ? A.flags[qq] : 0;
    int myBest = -1;
    const unsigned anyMask = __ballot_sync(FULL, myCnt > 0);
#pragma unroll 1
    for (int g = 0; g < 32; g += 8) {
      if (((anyMask >> g) & 0xffu) == 0) continue;   // warp-uniform
      uint32_t c[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int cnt = __shfl_sync(FULL, myCnt, g + u), ofs = __shfl_sync(FULL, myOfs, g + u);
        c[u] = lane < cnt ? A.cand[ofs + lane] : 0xffffffffu;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int cnt = __shfl_sync(FULL, myCnt, g + u);
        if (cnt == 0) continue;                      // warp-uniform
        const int ofs = __shfl_sync(FULL, myOfs, g + u);
        const int fl = __shfl_sync(FULL, myFlag, g + u);
        // two smallest (dist, position) keys among non-blocked candidates
        unsigned k1 = 0xffffffffu, k2 = 0xffffffffu;   // key = dist<<20 | position
        for (int b0 = 0; b0 < cnt; b0 += 32) {
          const int i = b0 + lane;
          unsigned key = 0xffffffffu;
          if (i < cnt) {
            const uint32_t cc = b0 == 0 ? c[u] : A.cand[ofs + i];
            if (!s_blocked[cc & 0xffff]) key = (((cc >> 16) & 0x1ff) << 20) | (unsigned)i;
          }
          const unsigned m1 = __reduce_min_sync(FULL, key);
          const unsigned m2 = __reduce_min_sync(FULL, key == m1 ? 0xffffffffu : key);
          // merge (m1,m2) into (k1,k2)
          if (m1 < k1) { k2 = min(k1, m2); k1 = m1; }
          else { k2 = min(k2, m1); }
        }
        int best = -1;
        if (k1 != 0xffffffffu) {
          const int bestDist = k1 >> 20;
          const int p1 = k1 & 0xfffff;
          const uint32_t c1 = p1 < 32 ? __shfl_sync(FULL, c[u], p1) : A.cand[ofs + p1];
          const int bestLevel = (c1 >> 25) & 0xf;
          int bestDist2 = 256, bestLevel2 = -1;
          if (k2 != 0xffffffffu) {
            bestDist2 = k2 >> 20;
            const int p2 = k2 & 0xfffff;
            const uint32_t c2 = p2 < 32 ? __shfl_sync(FULL, c[u], p2) : A.cand[ofs + p2];
            bestLevel2 = (c2 >> 25) & 0xf;
          }
          if (bestDist <= TH_HIGH) {
            const bool reject = bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(A.nnratio, (float)bestDist2);
            if (!reject) {
              best = c1 & 0xffff;
              __syncwarp();                            // every lane's s_blocked reads of this query are done
              if (lane == 0) s_blocked[best] = (fl & 2) ? 1 : 0;
              ++n;
              __syncwarp();
            }
          }
        }
        if (lane == g + u) myBest = best;
      }
    }
    if (qq < nq) A.bestIdx[qq] = myBest;
  }
  if (lane == 0) *A.nmatches = n;
}

// ------------------------------------------------------------------------------------
// K8b  SearchByProjection(Cur, Last)
// ------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) sbp_frame_score_kernel(const FrameDev* frames, const SbpFrameArgs* args) {
  SbpFrameArgs A = args[blockIdx.y];
  const FrameDev F = load_frame(frames, blockIdx.y);
  const int nq = A.nqDev ? *A.nqDev : A.nq;
  if (A.TcDev)
    for (int i = 0; i < 12; ++i) A.Tc[i] = A.TcDev[i];
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= nq) return;
  if (lane == 0) { A.candCnt[q] = 0; A.candOfs[q] = 0; }
  if (!(A.flags[q] & 1)) return;
  const float X = A.xw[3 * q], Y = A.xw[3 * q + 1], Z = A.xw[3 * q + 2];
  // x3Dc = Rcw*x3Dw + tcw in fp32, fixed left-to-right order
  const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A.Tc[0], X), __fmul_rn(A.Tc[1], Y)), __fmul_rn(A.Tc[2], Z)), A.Tc[3]);
  const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A.Tc[4], X), __fmul_rn(A.Tc[5], Y)), __fmul_rn(A.Tc[6], Z)), A.Tc[7]);
  const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A.Tc[8], X), __fmul_rn(A.Tc[9], Y)), __fmul_rn(A.Tc[10], Z)), A.Tc[11]);
  const float invzc = (float)(1.0 / (double)zc);
  if (invzc < 0) return;
  const float u = __fadd_rn(__fdiv_rn(__fmul_rn(A.fx, xc), zc), A.cx);
  const float v = __fadd_rn(__fdiv_rn(__fmul_rn(A.fy, yc), zc), A.cy);
  if (u < F.minX || u > F.maxX) return;
  if (v < F.minY || v > F.maxY) return;
  const int oct = A.octave[q];
  const float radius = __fmul_rn(A.th, A.scaleFactors[oct]);
  int minL, maxL;
  if (A.mode == 1) { minL = oct; maxL = -1; }
  else if (A.mode == 2) { minL = 0; maxL = oct; }
  else { minL = oct - 1; maxL = oct + 1; }
  const float ur = __fsub_rn(u, __fmul_rn(A.bf, invzc));
  auto gate = [&](int idx) {
    if (F.uright) {
      const float k = F.uright[idx];
      if (k > 0 && fabsf(__fsub_rn(ur, k)) > radius) return false;
    }
    return true;
  };
  int cache[SBP_CACHE], nIt;
  int cnt = walk_cached(F, u, v, radius, minL, maxL, gate, cache, nIt);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (cnt == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(A.total, cnt);
  base = __shfl_sync(0xffffffffu, base, 0);
  if (base + cnt > A.candCap) {
    if (lane == 0) atomicExch(A.err, 1);
    return;
  }
  const uint4* dq = reinterpret_cast<const uint4*>(A.mpDesc + 32 * (size_t)q);
  const CellRange c = cell_range(F, u, v, radius);
  int pos = 0;
  for (int it = 0; it < nIt; ++it) {                     // replay of the cached walk (nIt = -1: walk again below)
    const int idx = cache[it];
    const bool ok = idx >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int d = hamming256(dq, reinterpret_cast<const uint4*>(F.desc + 32 * (size_t)idx));
      A.cand[base + pos + __popc(m & ((1u << lane) - 1))] = (uint32_t)idx | ((uint32_t)d << 16);
    }
    pos += __popc(m);
  }
  for (int ix = c.x0; nIt < 0 && ix <= c.x1; ++ix) {
    const int beg = F.cellStart[ix * ORBX_GRID_ROWS + c.y0], end = F.cellStart[ix * ORBX_GRID_ROWS + c.y1 + 1];
    for (int b0 = beg; b0 < end; b0 += 32) {
      const int i = b0 + lane;
      int idx = -1;
      bool ok = false;
      if (i < end) {
        idx = F.cellIdx[i];
        ok = in_window(F.kps[idx], u, v, radius, minL, maxL) && gate(idx);
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int d = hamming256(dq, reinterpret_cast<const uint4*>(F.desc + 32 * (size_t)idx));
        A.cand[base + pos + __popc(m & ((1u << lane) - 1))] = (uint32_t)idx | ((uint32_t)d << 16);
      }
      pos += __popc(m);
    }
  }
  if (lane == 0) { A.candCnt[q] = cnt; A.candOfs[q] = base; }
}

__device__ __forceinline__ int rot_bin(float a, float b) {
  const float factor = 1.0f / HISTO_LENGTH;
  float rot = __fsub_rn(a, b);
  if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
  int bin = round_haz(__fmul_rn(rot, factor));
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

// ComputeThreeMaxima (src/ORBmatcher.cc:2654-2695) on bin counts; serial, 30 bins
__device__ void three_maxima(const int* h, int& i1, int& i2, int& i3) {
  int m1 = 0, m2 = 0, m3 = 0;
  i1 = i2 = i3 = -1;
  for (int i = 0; i < HISTO_LENGTH; ++i) {
    const int s = h[i];
    if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
    else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
    else if (s > m3) { m3 = s; i3 = i; }
  }
  if ((float)m2 < __fmul_rn(0.1f, (float)m1)) { i2 = -1; i3 = -1; }
  else if ((float)m3 < __fmul_rn(0.1f, (float)m1)) { i3 = -1; }
}

__global__ void __launch_bounds__(32) sbp_frame_resolve_kernel(const FrameDev* frames, const SbpFrameArgs* args) {
  extern __shared__ uint8_t s_blocked[];
  __shared__ int s_hist[HISTO_LENGTH];
  __shared__ int s_keep[3];
  const SbpFrameArgs A = args[blockIdx.x];
  const FrameDev F = load_frame(frames, blockIdx.x);
  const int nq = A.nqDev ? *A.nqDev : A.nq;
  const int lane = threadIdx.x;
  for (int i = lane; i < F.n; i += 32) { s_blocked[i] = A.curBlocked ? A.curBlocked[i] : 0; A.curMatch[i] = -1; }
  if (lane < HISTO_LENGTH) s_hist[lane] = 0;
  __syncwarp();
  // serial replay with coalesced per-32-query metadata and 8-query candidate prefetch (see sbp_map_resolve_kernel)
  const unsigned FULL = 0xffffffffu;
  int n = 0;
  for (int q0 = 0; q0 < nq; q0 += 32) {
    const int qq = q0 + lane;
    const int myCnt = qq < nq ? A.candCnt[qq] : 0;
    const int myOfs = qq < nq ? A.candOfs[qq] : 0;
    const int myFlag = qq < nq ? A.flags[qq] : 0;
    int myBest = -1;
    const unsigned anyMask = __ballot_sync(FULL, myCnt > 0);
#pragma unroll 1
    for (int g = 0; g < 32; g += 8) {
      if (((anyMask >> g) & 0xffu) == 0) continue;   // warp-uniform
      uint32_t c[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int cnt = __shfl_sync(FULL, myCnt, g + u), ofs = __shfl_sync(FULL, myOfs, g + u);
        c[u] = lane < cnt ? A.cand[ofs + lane] : 0xffffffffu;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int cnt = __shfl_sync(FULL, myCnt, g + u);
        if (cnt == 0) continue;                      // warp-uniform
        const int ofs = __shfl_sync(FULL, myOfs, g + u);
        const int fl = __shfl_sync(FULL, myFlag, g + u);
        unsigned k1 = 0xffffffffu;
        for (int b0 = 0; b0 < cnt; b0 += 32) {
          const int i = b0 + lane;
          unsigned key = 0xffffffffu;
          if (i < cnt) {
            const uint32_t cc = b0 == 0 ? c[u] : A.cand[ofs + i];
            if (!s_blocked[cc & 0xffff]) key = (((cc >> 16) & 0x1ff) << 20) | (unsigned)i;
          }
          k1 = min(k1, __reduce_min_sync(FULL, key));
        }
        int best = -1;
        if (k1 != 0xffffffffu && (int)(k1 >> 20) <= TH_HIGH) {
          const int p1 = k1 & 0xfffff;
          const uint32_t c1 = p1 < 32 ? __shfl_sync(FULL, c[u], p1) : A.cand[ofs + p1];
          best = c1 & 0xffff;
          __syncwarp();                                // every lane's s_blocked reads of this query are done
          if (lane == 0) {
            s_blocked[best] = (fl & 2) ? 1 : 0;
            A.curMatch[best] = q0 + g + u;
          }
          ++n;
          __syncwarp();
        }
        if (lane == g + u) myBest = best;
      }
    }
    if (qq < nq) { A.matchIdx[qq] = myBest; A.kept[qq] = myBest >= 0; }
    // rotation histogram of the accepted matches (counts only: order-independent)
    if (A.checkOri && myBest >= 0) atomicAdd(&s_hist[rot_bin(A.angle[qq], F.kps[myBest].angle)], 1);
  }
  __syncwarp();
  if (A.checkOri) {
    if (lane == 0) { int i1, i2, i3; three_maxima(s_hist, i1, i2, i3); s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3; }
    __syncwarp();
    const int i1 = s_keep[0], i2 = s_keep[1], i3 = s_keep[2];
    int removed = 0;
    for (int q = lane; q < nq; q += 32) {
      const int idx = A.matchIdx[q];
      if (idx < 0) continue;
      const int bin = rot_bin(A.angle[q], F.kps[idx].angle);
      if (bin != i1 && bin != i2 && bin != i3) {
        A.curMatch[idx] = -1;   // mvpMapPoints[idx] = NULL (even if a later query re-assigned it)
        A.kept[q] = 0;
        ++removed;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    n -= removed;
  }
  if (lane == 0) *A.nmatches = n;
}

// ------------------------------------------------------------------------------------
// K9  ComputeStereoMatches: warp per left keypoint (row-band Hamming + 11x11 SAD + parabola),
//     then one CTA per frame for the median SAD rejection.
// ------------------------------------------------------------------------------------

// Right keypoints are staged once per CTA in shared memory as compact records (x, row band, octave), so the row-band
// scan of every left keypoint (the reference's vRowIndices lookup, src/Frame.cc:972-982,1011-1038) reads shared memory
// instead of chasing kpR[i] -> scale[octave] through global memory for each of the nL x nR pairs.
#define STEREO_NT 256
#define STEREO_KPW 4          // left keypoints per warp
#define STEREO_CHUNK 2048     // right keypoints staged per pass

// sub-pixel refinement of one accepted descriptor match: 11x11 SAD over 11 shifts + parabola (src/Frame.cc:1041-1113)
__device__ __forceinline__ void stereo_refine(const StereoArgs& A, int iL, const orbx_keypoint& kl, unsigned key, int lane,
                                              float minD, float maxD) {
  if (key == 0xffffffffu) return;
  const int bestDist = key >> 16, bestIdxR = key & 0xffff;
  if (!(bestDist < TH_HIGH)) return;                 // bestDist starts at TH_HIGH: strict <
  if (!(bestDist < (TH_HIGH + TH_LOW) / 2)) return;
  const float uR0 = A.kpR[bestIdxR].x;
  const int oct = kl.octave;
  const float sf = A.invScale[oct];
  const float scaleduL = roundf(__fmul_rn(kl.x, sf)), scaledvL = roundf(__fmul_rn(kl.y, sf)), scaleduR0 = roundf(__fmul_rn(uR0, sf));
  const int w = 5, L = 5;
  const int W = A.lw[oct], H = A.lh[oct];
  const int cxL = (int)scaleduL, cy = (int)scaledvL, cxR = (int)scaleduR0;
  if (cy - w < 0 || cy + w >= H || cxL - w < 0 || cxL + w >= W) return;
  const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
  if (iniu < 0 || endu >= W) return;
  if (cxR - L - w < 0) return;
  const uint8_t* IL = A.pyrL[oct];
  const uint8_t* IR = A.pyrR[oct];
  const int pL = A.pitchL[oct], pR = A.pitchR[oct];
  const int cL = __ldg(IL + (size_t)cy * pL + cxL);
  // each lane owns up to 4 of the 121 patch pixels
  int la[4], ldy[4], ldx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = lane + 32 * k;
    ldy[k] = p / 11 - w;
    ldx[k] = p % 11 - w;
    la[k] = p < 121 ? (int)__ldg(IL + (size_t)(cy + ldy[k]) * pL + cxL + ldx[k]) - cL : 0;
  }
  int sads[11];
#pragma unroll
  for (int inc = -L; inc <= L; ++inc) {
    const int cR = __ldg(IR + (size_t)cy * pR + cxR + inc);
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int p = lane + 32 * k;
      if (p < 121) s += abs(la[k] - ((int)__ldg(IR + (size_t)(cy + ldy[k]) * pR + cxR + inc + ldx[k]) - cR));
    }
    sads[inc + L] = __reduce_add_sync(0xffffffffu, s);
  }
  if (lane != 0) return;
  int bestSad = 0x7fffffff, bestinc = 0;
#pragma unroll
  for (int i = 0; i < 11; ++i)
    if (sads[i] < bestSad) { bestSad = sads[i]; bestinc = i - L; }
  if (bestinc == -L || bestinc == L) return;
  float d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
  for (int i = 1; i < 10; ++i)
    if (i == bestinc + L) { d1 = (float)sads[i - 1]; d2 = (float)sads[i]; d3 = (float)sads[i + 1]; }
  const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
  if (deltaR < -1 || deltaR > 1) return;
  float bestuR = __fmul_rn(A.scale[oct], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
  float disparity = __fsub_rn(kl.x, bestuR);
  if (disparity >= minD && disparity < maxD) {
    if (disparity <= 0) {
      disparity = (float)0.01;
      bestuR = (float)((double)kl.x - 0.01);
    }
    A.depth[iL] = __fdiv_rn(A.bf, disparity);
    A.uright[iL] = bestuR;
    A.sad[iL] = bestSad;
  }
}

// Row index of the right keypoints: a counting sort of their ids by floor(y) (one CTA per frame).  The reference buckets
// every right keypoint into all rows of its band (vRowIndices, src/Frame.cc:965-982) and scans the bucket of the left
// keypoint's row; the band is at most +-(2 * scale[nlevels-1] + 1) rows wide, so scanning the rows v-W .. v+W of this index
// and re-testing the band visits a superset of that bucket.  The order inside a bucket does not matter: the minimum is
// taken over (distance << 16 | iR), i.e. the first minimum in ascending iR like the reference's strict "<".
__global__ void __launch_bounds__(256) stereo_rows_kernel(const StereoArgs* __restrict__ args) {
  __shared__ int s_cnt[ORBX_STEREO_MAX_ROWS + 1];
  __shared__ int s_warp[9];
  const StereoArgs& A = args[blockIdx.x];
  if (!A.sortIdx || !A.rowStart) return;
  const int nR = A.nRDev ? *A.nRDev : A.nR, nRows = A.lh[0], tid = threadIdx.x;
  if (nRows > ORBX_STEREO_MAX_ROWS) {
    if (tid == 0) A.rowStart[0] = -1;                 // not built: the matcher scans every right keypoint
    return;
  }
  for (int r = tid; r <= nRows; r += 256) s_cnt[r] = 0;
  __syncthreads();
  for (int i = tid; i < nR; i += 256) {
    const int r = min(max((int)floorf(A.kpR[i].y), 0), nRows - 1);
    atomicAdd(&s_cnt[r], 1);
  }
  __syncthreads();
  // exclusive scan of s_cnt[0..nRows) by 256 threads
  const int per = (nRows + 255) / 256, beg = min(tid * per, nRows), end = min(beg + per, nRows);
  int sum = 0;
  for (int r = beg; r < end; ++r) sum += s_cnt[r];
  int incl = sum;
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int w = 0; w < 8; ++w) { const int t = s_warp[w]; s_warp[w] = run; run += t; }
  }
  __syncthreads();
  int run = s_warp[wid] + incl - sum;
  for (int r = beg; r < end; ++r) {
    const int c = s_cnt[r];
    s_cnt[r] = run;
    A.rowStart[r] = run;
    run += c;
  }
  if (tid == 0) A.rowStart[nRows] = nR;
  __syncthreads();
  for (int i = tid; i < nR; i += 256) {
    const int r = min(max((int)floorf(A.kpR[i].y), 0), nRows - 1);
    A.sortIdx[atomicAdd(&s_cnt[r], 1)] = (uint16_t)i;
  }
}

__global__ void __launch_bounds__(STEREO_NT) stereo_match_kernel(const StereoArgs* __restrict__ args) {
  __shared__ float s_x[STEREO_CHUNK];
  __shared__ int s_rows[STEREO_CHUNK];      // minr (low 16, signed) | maxr << 16
  __shared__ uint8_t s_oct[STEREO_CHUNK];
  const StereoArgs& A = args[blockIdx.y];
  const int nL = A.nLDev ? *A.nLDev : A.nL, nR = A.nRDev ? *A.nRDev : A.nR;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if ((int)blockIdx.x * (STEREO_NT / 32) * STEREO_KPW >= nL) return;     // whole CTA out of range
  const int iL0 = ((int)blockIdx.x * (STEREO_NT / 32) + wid) * STEREO_KPW;
  const int nRows = A.lh[0];
  const float minZ = A.b, minD = 0.f, maxD = __fdiv_rn(A.bf, minZ);
  orbx_keypoint kl[STEREO_KPW];
  bool valid[STEREO_KPW];
  unsigned key[STEREO_KPW];   // dist<<16 | iR : first minimum in ascending iR
  int row[STEREO_KPW], octL[STEREO_KPW];
  float minU[STEREO_KPW], maxU[STEREO_KPW];
#pragma unroll
  for (int k = 0; k < STEREO_KPW; ++k) {
    const int iL = iL0 + k;
    valid[k] = iL < nL;
    key[k] = 0xffffffffu;
    row[k] = 0; octL[k] = 0; minU[k] = maxU[k] = 0.f;
    if (valid[k]) {
      if (lane == 0) { A.uright[iL] = -1.0f; A.depth[iL] = -1.0f; A.sad[iL] = -1; }
      kl[k] = A.kpL[iL];
      row[k] = (int)kl[k].y;
      minU[k] = __fsub_rn(kl[k].x, maxD);
      maxU[k] = __fsub_rn(kl[k].x, minD);
      valid[k] = row[k] >= 0 && row[k] < nRows && !(maxU[k] < 0);
      octL[k] = kl[k].octave;
    }
    if (!valid[k]) row[k] = -32768;
  }
  const bool indexed = A.sortIdx && A.rowStart && A.rowStart[0] >= 0;
  if (indexed) {
    // candidates of a left keypoint in row v: the right keypoints of rows v-W .. v+W of the index, band re-tested
    const int W = (int)ceilf(__fmul_rn(2.0f, A.scale[A.nlevels - 1])) + 1;
#pragma unroll
    for (int k = 0; k < STEREO_KPW; ++k) {
      if (!valid[k]) continue;                          // warp-uniform
      const int lo = A.rowStart[max(row[k] - W, 0)], hi = A.rowStart[min(row[k] + W + 1, nRows)];
      for (int i = lo + lane; i < hi; i += 32) {
        const int iR = A.sortIdx[i];
        const orbx_keypoint kr = A.kpR[iR];
        const float r = __fmul_rn(2.0f, A.scale[kr.octave]);
        const int maxr = (int)ceilf(__fadd_rn(kr.y, r)), minr = (int)floorf(__fsub_rn(kr.y, r));
        if (row[k] >= minr && row[k] <= maxr && kr.octave >= octL[k] - 1 && kr.octave <= octL[k] + 1 && kr.x >= minU[k] &&
            kr.x <= maxU[k]) {
          const int d = hamming256(reinterpret_cast<const uint4*>(A.descL + 32 * (size_t)(iL0 + k)),
                                   reinterpret_cast<const uint4*>(A.descR + 32 * (size_t)iR));
          key[k] = min(key[k], ((unsigned)d << 16) | (unsigned)iR);
        }
      }
    }
  }
  for (int c0 = 0; c0 < (indexed ? 0 : nR); c0 += STEREO_CHUNK) {
    const int n = min(STEREO_CHUNK, nR - c0);
    __syncthreads();                                  // the previous chunk has been consumed
    for (int i = tid; i < n; i += STEREO_NT) {
      const orbx_keypoint kr = A.kpR[c0 + i];
      const float r = __fmul_rn(2.0f, A.scale[kr.octave]);
      const int maxr = (int)ceilf(__fadd_rn(kr.y, r)), minr = (int)floorf(__fsub_rn(kr.y, r));
      s_x[i] = kr.x;
      s_rows[i] = (minr & 0xffff) | (maxr << 16);
      s_oct[i] = (uint8_t)kr.octave;
    }
    __syncthreads();
    // every staged record is tested against the warp's 4 left keypoints at once (one set of shared-memory loads)
    for (int i = lane; i < n; i += 32) {
      const int rows = s_rows[i];
      const int minr = (int)(short)(rows & 0xffff), maxr = rows >> 16;
      const int octR = s_oct[i];
      const float x = s_x[i];
#pragma unroll
      for (int k = 0; k < STEREO_KPW; ++k) {
        // (an invalid left keypoint has row = -32768: outside every band)
        if (row[k] >= minr && row[k] <= maxr && octR >= octL[k] - 1 && octR <= octL[k] + 1 && x >= minU[k] && x <= maxU[k]) {
          const int d = hamming256(reinterpret_cast<const uint4*>(A.descL + 32 * (size_t)(iL0 + k)),
                                   reinterpret_cast<const uint4*>(A.descR + 32 * (size_t)(c0 + i)));
          key[k] = min(key[k], ((unsigned)d << 16) | (unsigned)(c0 + i));
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < STEREO_KPW; ++k) {
    if (!valid[k]) continue;
    const unsigned best = __reduce_min_sync(0xffffffffu, key[k]);
    stereo_refine(A, iL0 + k, kl[k], best, lane, minD, maxD);
  }
}

// median of the accepted SADs (element size/2 of the sorted list) -> reject sad >= 1.5*1.4*median (src/Frame.cc:1119-1131).
// A SAD of 121 int16 differences is < 2^16, so the rank-size/2 element is found by a two-level 256-bin radix select.
__global__ void __launch_bounds__(256) stereo_median_kernel(const StereoArgs* __restrict__ args) {
  const StereoArgs& A = args[blockIdx.x];
  const int nL = A.nLDev ? *A.nLDev : A.nL;
  const int* sad = A.sad;
  float* uright = A.uright;
  float* depth = A.depth;
  __shared__ int s_hist[256];
  __shared__ int s_m, s_bin, s_before, s_med;
  const int tid = threadIdx.x;
  s_hist[tid] = 0;
  if (tid == 0) { s_m = 0; s_med = -1; }
  __syncthreads();
  int local = 0;
  for (int i = tid; i < nL; i += 256) {
    const int d = sad[i];
    if (d >= 0) { ++local; atomicAdd(&s_hist[min(d >> 8, 255)], 1); }
  }
  if (local) atomicAdd(&s_m, local);
  __syncthreads();
  const int M = s_m;
  if (M == 0) return;
  const int target = M / 2;
  if (tid == 0) {
    int cum = 0, bin = 0;
    for (; bin < 255; ++bin) {
      if (cum + s_hist[bin] > target) break;
      cum += s_hist[bin];
    }
    s_bin = bin;
    s_before = cum;
  }
  __syncthreads();
  const int bin = s_bin;
  __syncthreads();
  s_hist[tid] = 0;
  __syncthreads();
  for (int i = tid; i < nL; i += 256) {
    const int d = sad[i];
    if (d >= 0 && min(d >> 8, 255) == bin) atomicAdd(&s_hist[d & 255], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int cum = s_before, lo = 0;
    for (; lo < 255; ++lo) {
      if (cum + s_hist[lo] > target) break;
      cum += s_hist[lo];
    }
    s_med = (bin << 8) | lo;
  }
  __syncthreads();
  const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)s_med);
  for (int i = tid; i < nL; i += 256) {
    const int d = sad[i];
    if (d >= 0 && !((float)d < thDist)) { uright[i] = -1.0f; depth[i] = -1.0f; }
  }
}

// ------------------------------------------------------------------------------------
// K10  SearchForTriangulation: warp per KF1 feature (in FeatureVector order)
// ------------------------------------------------------------------------------------
struct TriArgs {
  int n1, n2;
  const uint8_t *has1, *has2;
  int nn1, nn2;
  const int *n1id, *n1off, *n1idx, *n2id, *n2off, *n2idx;
  float F12[9];
  float epx, epy;
  const float* sigma2;
  const float* scaleFactors;
  int onlyStereo, coarse, checkOri;
  int* match12;
  int* nmatches;
};

__device__ __forceinline__ void tri_match_body(const FrameDev& K1, const FrameDev& K2, const TriArgs& A) {
  const int p1 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (A.nn1 == 0 || A.nn2 == 0 || p1 >= A.n1off[A.nn1]) return;
  // node of position p1 (upper_bound on offsets), then the same node id in KF2
  int lo = 0, hi = A.nn1;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (A.n1off[mid] <= p1) lo = mid; else hi = mid; }
  const int node = A.n1id[lo];
  int l2 = 0, h2 = A.nn2;
  while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (A.n2id[mid] < node) l2 = mid + 1; else h2 = mid; }
  if (l2 >= A.nn2 || A.n2id[l2] != node) return;
  const int idx1 = A.n1idx[p1];
  if (A.has1[idx1]) return;
  const bool bStereo1 = K1.uright && K1.uright[idx1] >= 0;
  if (A.onlyStereo && !bStereo1) return;
  const orbx_keypoint kp1 = K1.kps[idx1];
  const uint4* d1 = reinterpret_cast<const uint4*>(K1.desc + 32 * (size_t)idx1);
  // epipolar line of kp1 in image 2: l = x1^T F12
  const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, A.F12[0]), __fmul_rn(kp1.y, A.F12[3])), A.F12[6]);
  const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, A.F12[1]), __fmul_rn(kp1.y, A.F12[4])), A.F12[7]);
  const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, A.F12[2]), __fmul_rn(kp1.y, A.F12[5])), A.F12[8]);
  const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
  // running-best semantics (dist <= best so far, gates before acceptance) == min dist, LAST on ties
  unsigned key = 0xffffffffu;   // dist<<20 | (0xfffff - position)
  const int b2 = A.n2off[l2], e2 = A.n2off[l2 + 1];
  for (int i2 = b2 + lane; i2 < e2; i2 += 32) {
    const int idx2 = A.n2idx[i2];
    if (A.has2[idx2]) continue;
    const bool bStereo2 = K2.uright && K2.uright[idx2] >= 0;
    if (A.onlyStereo && !bStereo2) continue;
    const int dist = hamming256(d1, reinterpret_cast<const uint4*>(K2.desc + 32 * (size_t)idx2));
    if (dist > TH_LOW) continue;
    const orbx_keypoint kp2 = K2.kps[idx2];
    if (!bStereo1 && !bStereo2) {
      const float dex = __fsub_rn(A.epx, kp2.x), dey = __fsub_rn(A.epy, kp2.y);
      if (__fadd_rn(__fmul_rn(dex, dex), __fmul_rn(dey, dey)) < __fmul_rn(100.f, A.scaleFactors[kp2.octave])) continue;
    }
    bool ok = A.coarse != 0;
    if (!ok && den != 0) {
      const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
      const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
      ok = (double)dsqr < 3.84 * (double)A.sigma2[kp2.octave];
    }
    if (ok) key = min(key, ((unsigned)dist << 20) | (unsigned)(0xfffff - (i2 - b2)));
  }
  key = __reduce_min_sync(0xffffffffu, key);
  if (lane == 0 && key != 0xffffffffu) A.match12[idx1] = A.n2idx[b2 + (0xfffff - (key & 0xfffff))];
}

__global__ void __launch_bounds__(128) tri_match_kernel(const FrameDev* frames, TriArgs A) { tri_match_body(frames[0], frames[1], A); }
// many (KF1, KF2) pairs per launch: blockIdx.y = pair, its frames at frames[2q], frames[2q+1]
__global__ void __launch_bounds__(128) tri_match_batch_kernel(const FrameDev* frames, const TriArgs* args) {
  const int q = blockIdx.y;
  tri_match_body(frames[2 * q], frames[2 * q + 1], args[q]);
}

__device__ __forceinline__ void tri_rot_filter_body(const FrameDev& K1, const FrameDev& K2, const TriArgs& A) {
  __shared__ int s_hist[HISTO_LENGTH];
  __shared__ int s_keep[3];
  __shared__ int s_n;
  if (threadIdx.x < HISTO_LENGTH) s_hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < A.n1; i += 256) {
    const int m = A.match12[i];
    if (m < 0) continue;
    ++local;
    if (A.checkOri) atomicAdd(&s_hist[rot_bin(K1.kps[i].angle, K2.kps[m].angle)], 1);
  }
  if (local) atomicAdd(&s_n, local);
  __syncthreads();
  if (A.checkOri) {
    if (threadIdx.x == 0) { int i1, i2, i3; three_maxima(s_hist, i1, i2, i3); s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3; }
    __syncthreads();
    int removed = 0;
    for (int i = threadIdx.x; i < A.n1; i += 256) {
      const int m = A.match12[i];
      if (m < 0) continue;
      const int bin = rot_bin(K1.kps[i].angle, K2.kps[m].angle);
      if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { A.match12[i] = -1; ++removed; }
    }
    if (removed) atomicSub(&s_n, removed);
    __syncthreads();
  }
  if (threadIdx.x == 0) *A.nmatches = s_n;
}
__global__ void __launch_bounds__(256) tri_rot_filter_kernel(const FrameDev* frames, TriArgs A) { tri_rot_filter_body(frames[0], frames[1], A); }
__global__ void __launch_bounds__(256) tri_rot_filter_batch_kernel(const FrameDev* frames, const TriArgs* args) {
  const int q = blockIdx.x;
  tri_rot_filter_body(frames[2 * q], frames[2 * q + 1], args[q]);
}

// =====================================================================================
// batched launchers
// =====================================================================================
// The ordered-replay kernels keep one byte per keypoint of the frame in dynamic shared memory.  Frames may hold up to
// 65535 keypoints (the candidate records carry 16-bit indices), i.e. more than the 48 KB a kernel gets without opting in:
// raise the limit once per device, and check every launch (a failed launch is not sticky -- without the check the entry
// point would return ORBX_OK with untouched output buffers).
static int resolve_smem_opt_in() {
  static std::mutex mu;
  static bool done[64] = {};
  int dev = 0;
  ORBX_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (!done[dev & 63]) {
    ORBX_CUDA(cudaFuncSetAttribute(sbp_frame_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64));
    ORBX_CUDA(cudaFuncSetAttribute(sbp_map_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64));
    done[dev & 63] = true;
  }
  return ORBX_OK;
}

int orbx_launch_stereo_batch(orbx_ctx* ctx, cudaStream_t st, const StereoArgs* dArgs, int S, int maxL) {
  stereo_rows_kernel<<<S, 256, 0, st>>>(dArgs);
  ORBX_LAUNCH(ctx);
  stereo_match_kernel<<<dim3(div_up(maxL, (STEREO_NT / 32) * STEREO_KPW), S), STEREO_NT, 0, st>>>(dArgs);
  ORBX_LAUNCH(ctx);
  stereo_median_kernel<<<S, 256, 0, st>>>(dArgs);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}
int orbx_launch_sbp_frame_batch(orbx_ctx* ctx, cudaStream_t st, const FrameDev* dF, const SbpFrameArgs* dA, int S, int maxQ,
                                int maxN) {
  if (maxN + 16 > 48 * 1024) { int rc = resolve_smem_opt_in(); if (rc != ORBX_OK) return rc; }
  sbp_frame_score_kernel<<<dim3(div_up(maxQ * 32, 128), S), 128, 0, st>>>(dF, dA);
  ORBX_LAUNCH(ctx);
  sbp_frame_resolve_kernel<<<S, 32, align_up((size_t)maxN + 16, 16), st>>>(dF, dA);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}
int orbx_launch_sbp_map_batch(orbx_ctx* ctx, cudaStream_t st, const FrameDev* dF, const SbpMapArgs* dA, int S, int maxQ,
                              int maxN) {
  if (maxN + 16 > 48 * 1024) { int rc = resolve_smem_opt_in(); if (rc != ORBX_OK) return rc; }
  sbp_map_score_kernel<<<dim3(div_up(maxQ * 32, 128), S), 128, 0, st>>>(dF, dA);
  ORBX_LAUNCH(ctx);
  sbp_map_resolve_kernel<<<S, 32, align_up((size_t)maxN + 16, 16), st>>>(dF, dA);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

// =====================================================================================
// host entry points
// =====================================================================================
struct orbx_ext;
int orbx_ext_pyramid_view(orbx_ext* e, int b, int* nlevels, const uint8_t** ptr, int* w, int* h, int* pitch, float* scale,
                          float* invScale, cudaStream_t* st);

// Epipole and fundamental matrix of a keyframe pair, fp32 with a fixed evaluation order (no FMA on the host: this TU's
// host code is compiled with -ffp-contract=off).  The reference rebuilds F12 for every candidate pair
// (src/CameraModels/Pinhole.cpp:155-160); it only depends on the two keyframes.
static void tri_epipolar_geometry(const orbx_camera* cam1, const orbx_camera* cam2, const float* R1w, const float* t1w,
                                  const float* R2w, const float* t2w, TriArgs& A) {
  float Cw[3], C2[3], R12[9], t12[3], Am[9], Bm[9];
  for (int i = 0; i < 3; ++i) Cw[i] = -(R1w[0 * 3 + i] * t1w[0] + R1w[1 * 3 + i] * t1w[1] + R1w[2 * 3 + i] * t1w[2]);
  for (int i = 0; i < 3; ++i) C2[i] = R2w[i * 3 + 0] * Cw[0] + R2w[i * 3 + 1] * Cw[1] + R2w[i * 3 + 2] * Cw[2] + t2w[i];
  A.epx = cam2->fx * C2[0] / C2[2] + cam2->cx;
  A.epy = cam2->fy * C2[1] / C2[2] + cam2->cy;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R12[i * 3 + j] = (float)((double)R1w[i * 3 + 0] * (double)R2w[j * 3 + 0] + (double)R1w[i * 3 + 1] * (double)R2w[j * 3 + 1] +
                               (double)R1w[i * 3 + 2] * (double)R2w[j * 3 + 2]);   // R1w*R2w.t(): gemm general path (double accumulation)
  for (int i = 0; i < 3; ++i) t12[i] = -(R12[i * 3 + 0] * t2w[0] + R12[i * 3 + 1] * t2w[1] + R12[i * 3 + 2] * t2w[2]) + t1w[i];
  const float tx[9] = {0, -t12[2], t12[1], t12[2], 0, -t12[0], -t12[1], t12[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Am[i * 3 + j] = tx[i * 3 + 0] * R12[0 * 3 + j] + tx[i * 3 + 1] * R12[1 * 3 + j] + tx[i * 3 + 2] * R12[2 * 3 + j];
  const float i1x = 1.0f / cam1->fx, i1y = 1.0f / cam1->fy, c1x = -cam1->cx * i1x, c1y = -cam1->cy * i1y;
  const float i2x = 1.0f / cam2->fx, i2y = 1.0f / cam2->fy, c2x = -cam2->cx * i2x, c2y = -cam2->cy * i2y;
  for (int j = 0; j < 3; ++j) {
    Bm[0 * 3 + j] = i1x * Am[0 * 3 + j];
    Bm[1 * 3 + j] = i1y * Am[1 * 3 + j];
    Bm[2 * 3 + j] = c1x * Am[0 * 3 + j] + c1y * Am[1 * 3 + j] + Am[2 * 3 + j];
  }
  for (int i = 0; i < 3; ++i) {
    A.F12[i * 3 + 0] = Bm[i * 3 + 0] * i2x;
    A.F12[i * 3 + 1] = Bm[i * 3 + 1] * i2y;
    A.F12[i * 3 + 2] = Bm[i * 3 + 0] * c2x + Bm[i * 3 + 1] * c2y + Bm[i * 3 + 2];
  }
}

static int candidate_capacity(int nq, int n) { return std::max(1 << 16, std::min(nq, 1 << 16) * 128 + n); }

extern "C" {

int orbx_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return d;
}

int orbx_features_in_area(orbx_ctx* ctx, const orbx_frame_desc* frame, int nq, const float* x, const float* y,
                          const float* r, const int32_t* minL, const int32_t* maxL, int32_t* out_idx, int cap,
                          int32_t* out_n) {
  if (!ctx || !frame || nq < 0 || cap < 1 || !x || !y || !r || !minL || !maxL || !out_idx || !out_n) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  FrameDev F;
  int rc = orbx_upload_frame(S, frame, &F);
  if (rc != ORBX_OK) return rc;
  FrameDev* dF = S.upload(&F, 1);
  float *dx = S.upload(x, nq), *dy = S.upload(y, nq), *dr = S.upload(r, nq);
  int *dmin = S.upload(minL, nq), *dmax = S.upload(maxL, nq);
  int* dout = S.alloc<int>((size_t)nq * cap);
  int* dn = S.alloc<int>(nq);
  if (S.failed) return ORBX_ECUDA;
  if (!F.gridBuilt) {
    rc = orbx_launch_grid_build(ctx, st, dF, 1);
    if (rc != ORBX_OK) return rc;
  }
  if (nq > 0) {
    ORBX_CUDA(cudaMemsetAsync(dout, 0xff, sizeof(int) * (size_t)nq * cap, st));   // slots past out_n[q] read back as -1
    features_in_area_kernel<<<div_up(nq * 32, 128), 128, 0, st>>>(dF, nq, dx, dy, dr, dmin, dmax, dout, cap, dn);
    ORBX_LAUNCH(ctx);
    ORBX_CUDA(cudaMemcpyAsync(out_idx, dout, sizeof(int) * (size_t)nq * cap, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaMemcpyAsync(out_n, dn, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
  }
  ORBX_CUDA(cudaStreamSynchronize(st));
  return ORBX_OK;
}

int orbx_search_by_projection_map(orbx_ctx* ctx, const orbx_frame_desc* frame, const uint8_t* kp_blocked, int nq,
                                  const float* proj_x, const float* proj_y, const float* proj_xr, const int32_t* level,
                                  const float* view_cos, const uint8_t* mp_desc, const uint8_t* flags, float th,
                                  float nnratio, const float* scale_factors, int nlevels, int32_t* best_idx,
                                  int32_t* nmatches) {
  if (!ctx || !frame || nq < 0 || !proj_x || !proj_y || !level || !view_cos || !mp_desc || !flags || !scale_factors ||
      nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !best_idx || !nmatches || frame->n > 65535)
    return ORBX_EINVAL;
  if (frame->uright && !proj_xr) return ORBX_EINVAL;
  for (int q = 0; q < nq; ++q)
    if ((flags[q] & 1) && (level[q] < 0 || level[q] >= nlevels)) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  FrameDev F;
  int rc = orbx_upload_frame(S, frame, &F);
  if (rc != ORBX_OK) return rc;
  FrameDev* dF = S.upload(&F, 1);
  SbpMapArgs A;
  A.nq = nq;
  A.projX = S.upload(proj_x, nq);
  A.projY = S.upload(proj_y, nq);
  A.projXR = proj_xr ? S.upload(proj_xr, nq) : nullptr;
  A.viewCos = S.upload(view_cos, nq);
  A.level = S.upload(level, nq);
  A.mpDesc = S.upload(mp_desc, (size_t)nq * 32);
  A.flags = S.upload(flags, nq);
  A.th = th;
  A.nnratio = nnratio;
  A.scaleFactors = S.upload(scale_factors, nlevels);
  A.candCap = candidate_capacity(nq, frame->n);
  A.candOfs = S.alloc<int>(nq);
  A.candCnt = S.alloc<int>(nq);
  A.cand = S.alloc<uint32_t>(A.candCap);
  int* misc = S.alloc<int>(4);
  A.total = misc;
  A.err = misc + 1;
  A.nmatches = misc + 2;
  A.kpBlocked = kp_blocked ? S.upload(kp_blocked, frame->n) : nullptr;
  A.bestIdx = S.alloc<int>(nq);
  if (S.failed) return ORBX_ECUDA;
  ORBX_CUDA(cudaMemsetAsync(misc, 0, 4 * sizeof(int), st));
  ORBX_CUDA(cudaMemsetAsync(A.bestIdx, 0xff, sizeof(int) * (size_t)std::max(nq, 1), st));   // -1: never hand back arena garbage
  if (!F.gridBuilt) {
    rc = orbx_launch_grid_build(ctx, st, dF, 1);
    if (rc != ORBX_OK) return rc;
  }
  A.nqDev = nullptr;
  SbpMapArgs* dA = S.upload(&A, 1);
  if (S.failed) return ORBX_ECUDA;
  if (nq > 0) {
    sbp_map_score_kernel<<<dim3(div_up(nq * 32, 128), 1), 128, 0, st>>>(dF, dA);
    ORBX_LAUNCH(ctx);
  }
  if (frame->n + 16 > 48 * 1024) { rc = resolve_smem_opt_in(); if (rc != ORBX_OK) return rc; }
  sbp_map_resolve_kernel<<<1, 32, align_up((size_t)frame->n + 16, 16), st>>>(dF, dA);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  int h[4];
  ORBX_CUDA(cudaMemcpyAsync(h, misc, sizeof h, cudaMemcpyDeviceToHost, st));
  if (nq > 0) ORBX_CUDA(cudaMemcpyAsync(best_idx, A.bestIdx, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  if (h[1]) {
    orbx_set_error("orbx_search_by_projection_map: candidate buffer overflow (%d > %d)", h[0], A.candCap);
    return ORBX_ECAP;
  }
  *nmatches = h[2];
  return ORBX_OK;
}

int orbx_search_by_projection_frame(orbx_ctx* ctx, const orbx_frame_desc* cur, const uint8_t* cur_blocked,
                                    const orbx_camera* cam, const float* Tcw_cur, const float* Tcw_last, int nq,
                                    const uint8_t* flags, const float* xw, const int32_t* octave, const float* angle,
                                    const uint8_t* mp_desc, float th, int bMono, int check_orientation,
                                    const float* scale_factors, int nlevels, int32_t* match_idx, uint8_t* kept,
                                    int32_t* cur_match, int32_t* nmatches) {
  if (!ctx || !cur || !cam || !Tcw_cur || !Tcw_last || nq < 0 || !flags || !xw || !octave || !angle || !mp_desc ||
      !scale_factors || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !match_idx || !kept || !cur_match || !nmatches ||
      cur->n > 65535)
    return ORBX_EINVAL;
  for (int q = 0; q < nq; ++q)
    if ((flags[q] & 1) && (octave[q] < 0 || octave[q] >= nlevels)) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  FrameDev F;
  int rc = orbx_upload_frame(S, cur, &F);
  if (rc != ORBX_OK) return rc;
  FrameDev* dF = S.upload(&F, 1);
  SbpFrameArgs A;
  A.nq = nq;
  A.flags = S.upload(flags, nq);
  A.xw = S.upload(xw, (size_t)nq * 3);
  A.octave = S.upload(octave, nq);
  A.angle = S.upload(angle, nq);
  A.mpDesc = S.upload(mp_desc, (size_t)nq * 32);
  for (int i = 0; i < 12; ++i) A.Tc[i] = Tcw_cur[i];
  A.fx = cam->fx; A.fy = cam->fy; A.cx = cam->cx; A.cy = cam->cy; A.bf = cam->bf;
  A.th = th;
  {
    // bForward / bBackward (src/ORBmatcher.cc:2258-2266): z of the current camera centre in the last frame
    const float* Tc = Tcw_cur;
    const float* Tl = Tcw_last;
    float twc[3];
    for (int i = 0; i < 3; ++i)   // -Rcw.t()*tcw: cv::gemm's general path (transposed operand): double accumulation, one rounding
      twc[i] = (float)(-((double)Tc[0 * 4 + i] * (double)Tc[3] + (double)Tc[1 * 4 + i] * (double)Tc[7] + (double)Tc[2 * 4 + i] * (double)Tc[11]));
    const float tlcz = Tl[8] * twc[0] + Tl[9] * twc[1] + Tl[10] * twc[2] + Tl[11];
    const bool fwd = tlcz > cam->b && !bMono, bwd = -tlcz > cam->b && !bMono;
    A.mode = fwd ? 1 : (bwd ? 2 : 0);
  }
  A.checkOri = check_orientation;
  A.scaleFactors = S.upload(scale_factors, nlevels);
  A.candCap = candidate_capacity(nq, cur->n);
  A.candOfs = S.alloc<int>(nq);
  A.candCnt = S.alloc<int>(nq);
  A.cand = S.alloc<uint32_t>(A.candCap);
  int* misc = S.alloc<int>(4);
  A.total = misc;
  A.err = misc + 1;
  A.nmatches = misc + 2;
  A.curBlocked = cur_blocked ? S.upload(cur_blocked, cur->n) : nullptr;
  A.matchIdx = S.alloc<int>(nq);
  A.kept = S.alloc<uint8_t>(nq);
  A.curMatch = S.alloc<int>(cur->n);
  if (S.failed) return ORBX_ECUDA;
  ORBX_CUDA(cudaMemsetAsync(A.matchIdx, 0xff, sizeof(int) * (size_t)std::max(nq, 1), st));   // -1 / 0: never hand back arena garbage
  ORBX_CUDA(cudaMemsetAsync(A.kept, 0, (size_t)std::max(nq, 1), st));
  ORBX_CUDA(cudaMemsetAsync(A.curMatch, 0xff, sizeof(int) * (size_t)std::max(cur->n, 1), st));
  if (S.failed) return ORBX_ECUDA;
  ORBX_CUDA(cudaMemsetAsync(misc, 0, 4 * sizeof(int), st));
  if (!F.gridBuilt) {
    rc = orbx_launch_grid_build(ctx, st, dF, 1);
    if (rc != ORBX_OK) return rc;
  }
  A.nqDev = nullptr;
  A.TcDev = nullptr;
  SbpFrameArgs* dA = S.upload(&A, 1);
  if (S.failed) return ORBX_ECUDA;
  if (nq > 0) {
    sbp_frame_score_kernel<<<dim3(div_up(nq * 32, 128), 1), 128, 0, st>>>(dF, dA);
    ORBX_LAUNCH(ctx);
  }
  if (cur->n + 16 > 48 * 1024) { rc = resolve_smem_opt_in(); if (rc != ORBX_OK) return rc; }
  sbp_frame_resolve_kernel<<<1, 32, align_up((size_t)cur->n + 16, 16), st>>>(dF, dA);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  int h[4];
  ORBX_CUDA(cudaMemcpyAsync(h, misc, sizeof h, cudaMemcpyDeviceToHost, st));
  if (nq > 0) {
    ORBX_CUDA(cudaMemcpyAsync(match_idx, A.matchIdx, sizeof(int) * nq, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaMemcpyAsync(kept, A.kept, nq, cudaMemcpyDeviceToHost, st));
  }
  if (cur->n > 0) ORBX_CUDA(cudaMemcpyAsync(cur_match, A.curMatch, sizeof(int) * cur->n, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  if (h[1]) {
    orbx_set_error("orbx_search_by_projection_frame: candidate buffer overflow (%d > %d)", h[0], A.candCap);
    return ORBX_ECAP;
  }
  *nmatches = h[2];
  return ORBX_OK;
}

int orbx_stereo_match(orbx_ctx* ctx, orbx_ext* extL, int bL, orbx_ext* extR, int bR, const orbx_keypoint* kpL,
                      const uint8_t* descL, int nL, const orbx_keypoint* kpR, const uint8_t* descR, int nR, float bf,
                      float b, float* uright, float* depth) {
  if (!ctx || !extL || !extR || nL < 0 || nR < 0 || nR > 65535 || !uright || !depth || !(b > 0)) return ORBX_EINVAL;
  if ((nL && (!kpL || !descL)) || (nR && (!kpR || !descR))) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  StereoArgs A;
  int nlv = 0, nlv2 = 0, w2[ORBX_MAX_LEVELS], h2[ORBX_MAX_LEVELS];
  float sc2[ORBX_MAX_LEVELS], isc2[ORBX_MAX_LEVELS];
  cudaStream_t stL, stR;
  int rc = orbx_ext_pyramid_view(extL, bL, &nlv, A.pyrL, A.lw, A.lh, A.pitchL, A.scale, A.invScale, &stL);
  if (rc != ORBX_OK) return rc;
  rc = orbx_ext_pyramid_view(extR, bR, &nlv2, A.pyrR, w2, h2, A.pitchR, sc2, isc2, &stR);
  if (rc != ORBX_OK) return rc;
  if (nlv != nlv2) return ORBX_EINVAL;
  for (int l = 0; l < nlv; ++l)
    if (w2[l] != A.lw[l] || h2[l] != A.lh[l]) {
      orbx_set_error("orbx_stereo_match: left/right pyramids differ in size");
      return ORBX_EINVAL;
    }
  for (int i = 0; i < nL; ++i)
    if (kpL[i].octave < 0 || kpL[i].octave >= nlv) return ORBX_EINVAL;
  for (int i = 0; i < nR; ++i)
    if (kpR[i].octave < 0 || kpR[i].octave >= nlv) return ORBX_EINVAL;
  // the pyramids were written on the extractors' streams
  ORBX_CUDA(cudaStreamSynchronize(stL));
  if (stR != stL) ORBX_CUDA(cudaStreamSynchronize(stR));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  A.nL = nL;
  A.nR = nR;
  A.nlevels = nlv;
  A.kpL = S.upload(kpL, nL);
  A.kpR = S.upload(kpR, nR);
  A.descL = S.upload(descL, (size_t)nL * 32);
  A.descR = S.upload(descR, (size_t)nR * 32);
  A.bf = bf;
  A.b = b;
  A.uright = S.alloc<float>(nL);
  A.depth = S.alloc<float>(nL);
  A.sad = S.alloc<int>(nL);
  A.sortIdx = S.alloc<uint16_t>(std::max(nR, 1));
  A.rowStart = S.alloc<int>(ORBX_STEREO_MAX_ROWS + 1);
  if (S.failed) return ORBX_ECUDA;
  A.nLDev = A.nRDev = nullptr;
  StereoArgs* dA = S.upload(&A, 1);
  if (S.failed) return ORBX_ECUDA;
  if (nL > 0) {
    stereo_rows_kernel<<<1, 256, 0, st>>>(dA);
    ORBX_LAUNCH(ctx);
    stereo_match_kernel<<<dim3(div_up(nL, (STEREO_NT / 32) * STEREO_KPW), 1), STEREO_NT, 0, st>>>(dA);
    ORBX_LAUNCH(ctx);
    stereo_median_kernel<<<1, 256, 0, st>>>(dA);
    ORBX_LAUNCH(ctx);
    ORBX_CUDA(cudaMemcpyAsync(uright, A.uright, sizeof(float) * nL, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaMemcpyAsync(depth, A.depth, sizeof(float) * nL, cudaMemcpyDeviceToHost, st));
  }
  ORBX_CUDA(cudaStreamSynchronize(st));
  return ORBX_OK;
}

int orbx_search_for_triangulation(orbx_ctx* ctx, const orbx_frame_desc* kf1, const orbx_frame_desc* kf2,
                                  const uint8_t* has_mp1, const uint8_t* has_mp2, int nn1, const int32_t* fv1_node,
                                  const int32_t* fv1_off, const int32_t* fv1_idx, int nn2, const int32_t* fv2_node,
                                  const int32_t* fv2_off, const int32_t* fv2_idx, const orbx_camera* cam1,
                                  const orbx_camera* cam2, const float* R1w, const float* t1w, const float* R2w,
                                  const float* t2w, const float* level_sigma2, const float* scale_factors, int nlevels,
                                  int only_stereo, int coarse, int check_orientation, int32_t* match12,
                                  int32_t* nmatches) {
  if (!ctx || !kf1 || !kf2 || !has_mp1 || !has_mp2 || nn1 < 0 || nn2 < 0 || !cam1 || !cam2 || !R1w || !t1w || !R2w ||
      !t2w || !level_sigma2 || !scale_factors || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !match12 || !nmatches)
    return ORBX_EINVAL;
  if ((nn1 && (!fv1_node || !fv1_off || !fv1_idx)) || (nn2 && (!fv2_node || !fv2_off || !fv2_idx))) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  FrameDev F[2];
  int rc = orbx_upload_frame(S, kf1, &F[0]);
  if (rc != ORBX_OK) return rc;
  rc = orbx_upload_frame(S, kf2, &F[1]);
  if (rc != ORBX_OK) return rc;
  FrameDev* dF = S.upload(F, 2);
  TriArgs A;
  A.n1 = kf1->n;
  A.n2 = kf2->n;
  A.has1 = S.upload(has_mp1, kf1->n);
  A.has2 = S.upload(has_mp2, kf2->n);
  A.nn1 = nn1;
  A.nn2 = nn2;
  const int tot1 = nn1 ? fv1_off[nn1] : 0, tot2 = nn2 ? fv2_off[nn2] : 0;
  static const int zero = 0;
  A.n1id = S.upload(nn1 ? fv1_node : &zero, std::max(nn1, 1));
  A.n1off = S.upload(nn1 ? fv1_off : &zero, nn1 + 1);
  A.n1idx = S.upload(tot1 ? fv1_idx : &zero, std::max(tot1, 1));
  A.n2id = S.upload(nn2 ? fv2_node : &zero, std::max(nn2, 1));
  A.n2off = S.upload(nn2 ? fv2_off : &zero, nn2 + 1);
  A.n2idx = S.upload(tot2 ? fv2_idx : &zero, std::max(tot2, 1));
  tri_epipolar_geometry(cam1, cam2, R1w, t1w, R2w, t2w, A);
  A.sigma2 = S.upload(level_sigma2, nlevels);
  A.scaleFactors = S.upload(scale_factors, nlevels);
  A.onlyStereo = only_stereo;
  A.coarse = coarse;
  A.checkOri = check_orientation;
  A.match12 = S.alloc<int>(kf1->n);
  A.nmatches = S.alloc<int>(1);
  if (S.failed) return ORBX_ECUDA;
  ORBX_CUDA(cudaMemsetAsync(A.match12, 0xff, sizeof(int) * std::max(kf1->n, 1), st));
  ORBX_CUDA(cudaMemsetAsync(A.nmatches, 0, sizeof(int), st));
  if (tot1 > 0 && nn2 > 0) {
    tri_match_kernel<<<div_up(tot1 * 32, 128), 128, 0, st>>>(dF, A);
    ORBX_LAUNCH(ctx);
  }
  tri_rot_filter_kernel<<<1, 256, 0, st>>>(dF, A);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  if (kf1->n > 0) ORBX_CUDA(cudaMemcpyAsync(match12, A.match12, sizeof(int) * kf1->n, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaMemcpyAsync(nmatches, A.nmatches, sizeof(int), cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  return ORBX_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// Many SearchForTriangulation calls per launch (the keyframe-rate half of the batched mode: LocalMapping::
// CreateNewMapPoints runs one call per covisible neighbour, src/LocalMapping.cc:501-628).  prepare() uploads the Q
// problems once into one device pool; run() only enqueues (two memsets + two launches) on the given stream, so a
// replay can overlap the per-frame chain; fetch() synchronises that stream and copies the matches out.
// ------------------------------------------------------------------------------------------------------------------
struct orbx_tri_batch {
  orbx_ctx* ctx = nullptr;
  int device = 0;               // kept separately: destroy() must not dereference a context that may be gone already
  int Q = 0, maxTot1 = 0;
  uint8_t* pool = nullptr;
  FrameDev* dFrames = nullptr;
  TriArgs* dArgs = nullptr;
  int* dMatch = nullptr;        // all problems' match12, back to back
  size_t matchInts = 0;
  int* dNm = nullptr;           // [Q]
  std::vector<int> matchOfs, n1;
  cudaStream_t lastStream = nullptr;
};

orbx_tri_batch* orbx_tri_batch_prepare(orbx_ctx* ctx, int Q, const orbx_tri_problem* pr, const float* level_sigma2,
                                       const float* scale_factors, int nlevels, int check_orientation) {
  if (!ctx || Q < 1 || !pr || !level_sigma2 || !scale_factors || nlevels < 1 || nlevels > ORBX_MAX_LEVELS) {
    orbx_set_error("orbx_tri_batch_prepare: invalid argument");
    return nullptr;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
  PoolBuilder B;
  const size_t oSig = B.add(level_sigma2, sizeof(float) * nlevels), oSc = B.add(scale_factors, sizeof(float) * nlevels);
  struct Ofs { size_t kps[2], desc[2], ur[2], has[2], id[2], off[2], idx[2]; };
  std::vector<Ofs> O(Q);
  std::vector<int> matchOfs(Q + 1, 0), n1(Q);
  int maxTot1 = 0;
  static const int zero = 0;
  for (int q = 0; q < Q; ++q) {
    const orbx_tri_problem& P = pr[q];
    const orbx_frame_desc* kf[2] = {P.kf1, P.kf2};
    const uint8_t* has[2] = {P.has_mp1, P.has_mp2};
    const int nn[2] = {P.nn1, P.nn2};
    const int32_t *id[2] = {P.fv1_node, P.fv2_node}, *off[2] = {P.fv1_off, P.fv2_off}, *idx[2] = {P.fv1_idx, P.fv2_idx};
    if (!P.kf1 || !P.kf2 || !P.has_mp1 || !P.has_mp2 || !P.cam1 || !P.cam2 || !P.R1w || !P.t1w || !P.R2w || !P.t2w || P.nn1 < 0 || P.nn2 < 0 ||
        (P.nn1 && (!P.fv1_node || !P.fv1_off || !P.fv1_idx)) || (P.nn2 && (!P.fv2_node || !P.fv2_off || !P.fv2_idx))) {
      orbx_set_error("orbx_tri_batch_prepare: problem %d has a NULL field", q);
      return nullptr;
    }
    for (int k = 0; k < 2; ++k) {
      const orbx_frame_desc* f = kf[k];
      if (f->n < 0 || (f->n > 0 && (!f->kps || !f->desc))) { orbx_set_error("orbx_tri_batch_prepare: invalid frame"); return nullptr; }
      O[q].kps[k] = B.add(f->kps, sizeof(orbx_keypoint) * (size_t)f->n);
      O[q].desc[k] = B.add(f->desc, (size_t)32 * f->n);
      O[q].ur[k] = f->uright ? B.add(f->uright, sizeof(float) * (size_t)f->n) : (size_t)-1;
      O[q].has[k] = B.add(has[k], (size_t)f->n);
      const int tot = nn[k] ? off[k][nn[k]] : 0;
      O[q].id[k] = B.add(nn[k] ? id[k] : &zero, sizeof(int) * (size_t)std::max(nn[k], 1));
      O[q].off[k] = B.add(nn[k] ? off[k] : &zero, sizeof(int) * (size_t)(nn[k] + 1));
      O[q].idx[k] = B.add(tot ? idx[k] : &zero, sizeof(int) * (size_t)std::max(tot, 1));
      if (k == 0) maxTot1 = std::max(maxTot1, tot);
    }
    n1[q] = P.kf1->n;
    matchOfs[q + 1] = matchOfs[q] + std::max(P.kf1->n, 1);
  }
  const size_t oMatch = B.reserve(sizeof(int) * (size_t)matchOfs[Q]), oNm = B.reserve(sizeof(int) * (size_t)Q);
  const size_t oFrames = B.reserve(sizeof(FrameDev) * 2 * (size_t)Q), oArgs = B.reserve(sizeof(TriArgs) * (size_t)Q);
  orbx_tri_batch* T = new orbx_tri_batch();
  T->ctx = ctx; T->device = ctx->device; T->Q = Q; T->maxTot1 = maxTot1; T->matchOfs = matchOfs; T->n1 = n1; T->matchInts = (size_t)matchOfs[Q];
  if (cudaMalloc(&T->pool, B.size()) != cudaSuccess) {
    orbx_set_error("orbx_tri_batch_prepare: cudaMalloc(%zu) failed", B.size());
    cudaGetLastError();
    delete T;
    return nullptr;
  }
  uint8_t* base = T->pool;
  std::vector<FrameDev> F(2 * (size_t)Q);
  std::vector<TriArgs> A(Q);
  for (int q = 0; q < Q; ++q) {
    const orbx_tri_problem& P = pr[q];
    const orbx_frame_desc* kf[2] = {P.kf1, P.kf2};
    for (int k = 0; k < 2; ++k) {
      FrameDev& f = F[2 * (size_t)q + k];
      memset(&f, 0, sizeof f);
      f.n = kf[k]->n;
      f.kps = (const orbx_keypoint*)(base + O[q].kps[k]);
      f.desc = base + O[q].desc[k];
      f.uright = O[q].ur[k] == (size_t)-1 ? nullptr : (const float*)(base + O[q].ur[k]);
      f.minX = kf[k]->min_x; f.minY = kf[k]->min_y; f.maxX = kf[k]->max_x; f.maxY = kf[k]->max_y;
    }
    TriArgs& a = A[q];
    a.n1 = P.kf1->n; a.n2 = P.kf2->n;
    a.has1 = base + O[q].has[0]; a.has2 = base + O[q].has[1];
    a.nn1 = P.nn1; a.nn2 = P.nn2;
    a.n1id = (const int*)(base + O[q].id[0]); a.n1off = (const int*)(base + O[q].off[0]); a.n1idx = (const int*)(base + O[q].idx[0]);
    a.n2id = (const int*)(base + O[q].id[1]); a.n2off = (const int*)(base + O[q].off[1]); a.n2idx = (const int*)(base + O[q].idx[1]);
    tri_epipolar_geometry(P.cam1, P.cam2, P.R1w, P.t1w, P.R2w, P.t2w, a);
    a.sigma2 = (const float*)(base + oSig);
    a.scaleFactors = (const float*)(base + oSc);
    a.onlyStereo = P.only_stereo; a.coarse = P.coarse; a.checkOri = check_orientation;
    a.match12 = (int*)(base + oMatch) + matchOfs[q];
    a.nmatches = (int*)(base + oNm) + q;
  }
  memcpy(B.h.data() + oFrames, F.data(), sizeof(FrameDev) * F.size());
  memcpy(B.h.data() + oArgs, A.data(), sizeof(TriArgs) * A.size());
  // (pageable H2D on the legacy stream returns once the data is staged: wait for the DMA, the plan runs on non-blocking streams)
  if (cudaMemcpy(T->pool, B.h.data(), B.size(), cudaMemcpyHostToDevice) != cudaSuccess || cudaStreamSynchronize(0) != cudaSuccess) {
    orbx_set_error("orbx_tri_batch_prepare: upload failed");
    cudaGetLastError();
    cudaFree(T->pool);
    delete T;
    return nullptr;
  }
  T->dFrames = (FrameDev*)(base + oFrames);
  T->dArgs = (TriArgs*)(base + oArgs);
  T->dMatch = (int*)(base + oMatch);
  T->dNm = (int*)(base + oNm);
  return T;
}

int orbx_tri_batch_run(orbx_tri_batch* T, void* cuda_stream) {
  if (!T) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(T->ctx->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : T->ctx->stream;
  ORBX_CUDA(cudaMemsetAsync(T->dMatch, 0xff, sizeof(int) * T->matchInts, st));
  ORBX_CUDA(cudaMemsetAsync(T->dNm, 0, sizeof(int) * (size_t)T->Q, st));
  if (T->maxTot1 > 0) {
    tri_match_batch_kernel<<<dim3(div_up(T->maxTot1 * 32, 128), T->Q), 128, 0, st>>>(T->dFrames, T->dArgs);
    ORBX_LAUNCH(T->ctx);
  }
  tri_rot_filter_batch_kernel<<<T->Q, 256, 0, st>>>(T->dFrames, T->dArgs);
  ORBX_LAUNCH(T->ctx);
  ORBX_CUDA(cudaGetLastError());
  T->lastStream = st;
  return ORBX_OK;
}

int orbx_tri_batch_fetch(orbx_tri_batch* T, orbx_tri_problem* pr) {
  if (!T || !pr) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(T->ctx->device));
  ORBX_CUDA(cudaStreamSynchronize(T->lastStream ? T->lastStream : T->ctx->stream));
  std::vector<int> m(T->matchInts), nm(T->Q);
  ORBX_CUDA(cudaMemcpy(m.data(), T->dMatch, sizeof(int) * T->matchInts, cudaMemcpyDeviceToHost));
  ORBX_CUDA(cudaMemcpy(nm.data(), T->dNm, sizeof(int) * (size_t)T->Q, cudaMemcpyDeviceToHost));
  for (int q = 0; q < T->Q; ++q) {
    if (pr[q].match12 && T->n1[q] > 0) memcpy(pr[q].match12, m.data() + T->matchOfs[q], sizeof(int) * (size_t)T->n1[q]);
    pr[q].nmatches = nm[q];
  }
  return ORBX_OK;
}

void orbx_tri_batch_destroy(orbx_tri_batch* T) {
  if (!T) return;
  cudaSetDevice(T->device);
  cudaDeviceSynchronize();   // the stream of the last run may belong to an object that is gone already
  cudaFree(T->pool);
  delete T;
}


// ------------------------------------------------------------------------------------------------------------------
// Device-resident Frame / KeyFrame.  A Frame crosses the boundary ONCE: keypoints, descriptors and
// mvuRight are copied to the device and the 64x48 grid (Frame::AssignFeaturesToGrid) is built; every later matcher call
// that is handed a descriptor over the same host arrays (SearchByProjection x2, SearchForTriangulation, Fuse,
// GetFeaturesInArea) finds the resident copy and skips its uploads and the grid build.  The caller keeps the host arrays
// unchanged while the handle lives (the reference's Frame is immutable in these fields after its constructor) and
// releases the handle when the Frame dies.
// ------------------------------------------------------------------------------------------------------------------
orbx_frame* orbx_frame_upload(orbx_ctx* ctx, const orbx_frame_desc* f) {
  if (!ctx || !f || f->n < 0 || (f->n > 0 && (!f->kps || !f->desc)) || !(f->max_x > f->min_x) || !(f->max_y > f->min_y) || f->n > 65535) {
    orbx_set_error("orbx_frame_upload: invalid frame descriptor");
    return nullptr;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
  PoolBuilder B;
  const size_t n = (size_t)f->n;
  const size_t oK = B.add(f->kps, sizeof(orbx_keypoint) * n), oD = B.add(f->desc, 32 * n);
  const size_t oU = f->uright ? B.add(f->uright, sizeof(float) * n) : (size_t)-1;
  const size_t oCs = B.reserve(sizeof(int) * (ORBX_NCELLS + 1)), oCi = B.reserve(sizeof(int) * n), oF = B.reserve(sizeof(FrameDev));
  orbx_frame* R = new orbx_frame();
  R->ctx = ctx;
  R->keyKps = f->kps; R->keyDesc = f->desc; R->keyUright = f->uright;
  R->n = f->n;
  R->bounds[0] = f->min_x; R->bounds[1] = f->min_y; R->bounds[2] = f->max_x; R->bounds[3] = f->max_y;
  if (cudaMalloc(&R->pool, B.size()) != cudaSuccess) {
    orbx_set_error("orbx_frame_upload: cudaMalloc(%zu) failed", B.size());
    cudaGetLastError();
    delete R;
    return nullptr;
  }
  uint8_t* base = R->pool;
  FrameDev& F = R->F;
  memset(&F, 0, sizeof F);
  F.n = f->n;
  F.kps = (const orbx_keypoint*)(base + oK);
  F.desc = base + oD;
  F.uright = oU == (size_t)-1 ? nullptr : (const float*)(base + oU);
  F.minX = f->min_x; F.minY = f->min_y; F.maxX = f->max_x; F.maxY = f->max_y;
  F.wInv = (float)ORBX_GRID_COLS / (f->max_x - f->min_x);
  F.hInv = (float)ORBX_GRID_ROWS / (f->max_y - f->min_y);
  F.cellStart = (int*)(base + oCs);
  F.cellIdx = (int*)(base + oCi);
  F.gridBuilt = 1;
  memcpy(B.h.data() + oF, &F, sizeof F);
  std::lock_guard<std::mutex> lock(ctx->apiMutex);
  bool ok = cudaMemcpyAsync(base, B.h.data(), B.h.size(), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
  ok = ok && orbx_launch_grid_build(ctx, ctx->stream, (const FrameDev*)(base + oF), 1) == ORBX_OK;
  ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;   // B.h is a local
  if (!ok) {
    orbx_set_error("orbx_frame_upload: upload failed");
    cudaGetLastError();
    cudaFree(R->pool);
    delete R;
    return nullptr;
  }
  ctx->residentFrames.push_back(R);
  return R;
}

void orbx_frame_release(orbx_frame* R) {
  if (!R) return;
  orbx_ctx* ctx = R->ctx;
  {
    std::lock_guard<std::mutex> lock(ctx->apiMutex);
    auto& v = ctx->residentFrames;
    v.erase(std::remove(v.begin(), v.end(), R), v.end());
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(R->pool);
  }
  delete R;
}

int orbx_frame_count(const orbx_ctx* ctx) { return ctx ? (int)ctx->residentFrames.size() : ORBX_EINVAL; }

}  // extern "C"
