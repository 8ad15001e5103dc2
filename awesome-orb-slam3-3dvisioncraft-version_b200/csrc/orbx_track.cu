// orbx_track.cu — many-stream tracking replay: the per-frame chain of Tracking::Track for S independent
// stereo streams, device-resident between stages.
//
//   Frame::Frame(stereo)            src/Frame.cc:90-170      -> extractor (2*S images), ComputeStereoMatches
//   Tracking::TrackWithMotionModel  src/Tracking.cc:2331-2433 -> SearchByProjection(Cur, Last, th), PoseOptimization,
//                                                                 outlier MapPoints dropped (:2402-2419)
//   Tracking::TrackLocalMap         src/Tracking.cc:2436-2480 -> SearchLocalPoints (:2848-2967: frustum test, points
//                                                                 already matched are skipped), SearchByProjection(F,
//                                                                 local points, th), PoseOptimization
// Two sources for the map each stream tracks against (no dataset offline):
//   * self-map (round 1, kept for continuity): the frame's own stereo points back-projected at the true pose stand in
//     for the last frame's / local map's MapPoints, with the frame's own descriptors (every true match has Hamming 0);
//   * given map (orbx_tracker_set_map): the caller supplies, per stream, the local map as flat arrays —
//     world positions, descriptors (bit-flipped copies + distractors), last-frame bookkeeping, distance-invariance
//     range and normals — as the shim would flatten Tracking's mvpLocalMapPoints; the local-map search then runs the
//     full Frame::isInFrustum test (src/Frame.cc:571-650) with MapPoint::PredictScale.
// With orbx_tracker_set_chain the motion-model prior of step t+1 is built on the device from the pose step t produced
// (Tracking::TrackWithMotionModel: mCurrentFrame.SetPose(mVelocity*mLastFrame.mTcw), src/Tracking.cc:2354).
// The glue kernels here (back-projection, edge gathering, frustum test) are the device form of that harness; the heavy
// kernels are the ones behind the single-frame C ABI.
#include <algorithm>
#include <cmath>
#include <vector>
#include "orbx_match.cuh"

struct orbx_ext;
uint8_t* orbx_ext_level0_storage(orbx_ext* e, size_t* bytes);   // orbx_extract.cu
int orbx_ext_pyramid_view(orbx_ext* e, int b, int* nlevels, const uint8_t** ptr, int* w, int* h, int* pitch, float* scale,
                          float* invScale, cudaStream_t* st);
int orbx_launch_pose_opt_slices(orbx_ctx* ctx, cudaStream_t st, int P, const int* d_start, const int* d_count, const float* d_xw,
                                const float* d_obs, const float* d_isg, const orbx_camera* cam, float* d_Tcw, uint8_t* d_outlier,
                                int* d_ninl, int* d_iters, double* d_scratch);
size_t orbx_inertial_args_bytes(int S);                          // orbx_optimize.cu
int orbx_launch_pose_inertial_slices(orbx_ctx* ctx, cudaStream_t st, const OrbxInertialSlices& I, void* d_args);

struct TrackDev {
  int S, cap;
  int ips;                    // images per stream: 2 = stereo (left, right), 1 = monocular
  int mcap;                   // stride of the map-indexed arrays (== cap in self-map mode)
  int useMap;                 // 1: map given by the caller (orbx_tracker_set_map)
  const int* nMap;            // [S] map points per stream (given map) or null
  const uint8_t* gMapFlags;   // [S][mcap] given map: bit0 in the local map && !isBad, bit1 Observations()>0
  const float *gMaxDist, *gMinDist, *gNormal;
  float logScale;
  int nlevels;
  float fx, fy, cx, cy, bf;
  float invSigma2[ORBX_MAX_LEVELS];
  const orbx_keypoint* kps;   // [2S][cap]
  const uint8_t* desc;        // [2S][cap][32]
  const int* n;               // [2S]
  float *uright, *depth;      // [S][cap]
  // "map points" = the frame's own stereo points
  uint8_t* mpFlags;           // [S][mcap] local map: bit0 has MapPoint, bit1 Observations()>0 (self-map mode)
  uint8_t* lastFlags;         // [S][mcap] subset of the map the *last frame* had tracked (self-map: 3 of every 5 points)
  const float* xw;            // [S][mcap][3]
  int* octave;                // [S][mcap] (self-map mode)
  float* angle;               // [S][mcap]
  float* xwOwn;               // self-map mode: the tracker's own storage behind xw
  // search 1 outputs
  int *matchIdx, *curMatch, *nm1;
  uint8_t* kept;
  // pose-opt edges (slice s = [s*cap, s*cap + count[s]))
  float *exw, *eobs, *eisg;
  int *ekp, *ecount, *estart, *kpEdge;
  uint8_t* eoutlier;
  uint8_t* eclose;            // pMP->mTrackDepth < 10 (given map: bit 2 of map_flags), read by the inertial optimisers
  int *ninl, *iters;
  // local-map search
  uint8_t *blocked, *mpTaken, *mapFlags;
  float *projX, *projY, *projXR, *viewCos;
  int *level, *bestIdx, *nm2, *kpMp;
  float *Ttrue, *Tprior, *T1, *T2;
  int* stats;
};

// K-glue 1: back-project stereo keypoints at the true pose: Xw = Rcw^T (Xc - tcw)
__global__ void __launch_bounds__(128) backproject_kernel(const TrackDev D) {
  const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.cap) return;
  const size_t o = (size_t)s * D.mcap + i;           // map-indexed arrays (keypoint i is map point i in this mode)
  const int n = D.n[D.ips * s];
  uint8_t fl = 0, lfl = 0;
  if (i < n) {
    const orbx_keypoint kp = D.kps[(size_t)(D.ips * s) * D.cap + i];
    const float z = D.depth[(size_t)s * D.cap + i];
    D.octave[o] = kp.octave;
    D.angle[o] = kp.angle;
    if (z > 0) {
      const float* T = D.Ttrue + 16 * s;
      const float xc = (kp.x - D.cx) * z / D.fx, yc = (kp.y - D.cy) * z / D.fy;
      const float dx = xc - T[3], dy = yc - T[7], dz = z - T[11];
      D.xwOwn[3 * o] = T[0] * dx + T[4] * dy + T[8] * dz;
      D.xwOwn[3 * o + 1] = T[1] * dx + T[5] * dy + T[9] * dz;
      D.xwOwn[3 * o + 2] = T[2] * dx + T[6] * dy + T[10] * dz;
      fl = 3;
      // the last frame had tracked ~60 % of the local map; the rest is only reachable through TrackLocalMap
      lfl = (((unsigned)i * 2654435761u) >> 16) % 5u < 3u ? 3 : 0;
    }
  }
  D.mpFlags[o] = fl;
  D.lastFlags[o] = lfl;
}

// K-glue 1b (monocular): no ComputeStereoMatches -- mvuRight = mvDepth = -1 for every keypoint (src/Frame.cc:330-331)
__global__ void __launch_bounds__(128) mono_fill_kernel(const TrackDev D) {
  const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.cap) return;
  const size_t o = (size_t)s * D.cap + i;
  D.uright[o] = -1.0f;
  D.depth[o] = -1.0f;
}

// K-glue 2: PoseOptimization's edge list: keypoints i (ascending) that hold a MapPoint (src/Optimizer.cc:961-1130)
// mode 0: MapPoint of keypoint i = curMatch[i]; mode 1: kpMp[i]
__global__ void __launch_bounds__(256) gather_edges_kernel(const TrackDev D, int mode) {
  const int s = blockIdx.x, tid = threadIdx.x;
  const int n = D.n[D.ips * s];
  const size_t base = (size_t)s * D.cap, mb = (size_t)s * D.mcap;
  const int* src = (mode == 0 ? D.curMatch : D.kpMp) + base;
  __shared__ int s_warp[9];
  __shared__ int s_run;
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + tid;
    const int q = i < n ? src[i] : -1;
    const int has = q >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, has);
    const int lane = tid & 31, wid = tid >> 5;
    if (lane == 0) s_warp[wid] = __popc(m);
    __syncthreads();
    int before = s_run;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    const int slot = before + __popc(m & ((1u << lane) - 1));
    if (i < n) D.kpEdge[base + i] = has ? slot : -1;
    if (has) {
      const orbx_keypoint kp = D.kps[(size_t)(D.ips * s) * D.cap + i];
      const size_t e = base + slot;
      D.exw[3 * e] = D.xw[3 * (mb + q)];
      D.exw[3 * e + 1] = D.xw[3 * (mb + q) + 1];
      D.exw[3 * e + 2] = D.xw[3 * (mb + q) + 2];
      D.eobs[3 * e] = kp.x;
      D.eobs[3 * e + 1] = kp.y;
      D.eobs[3 * e + 2] = D.uright[base + i];
      D.eisg[e] = D.invSigma2[kp.octave];
      D.eclose[e] = D.useMap ? ((D.gMapFlags[mb + q] >> 2) & 1) : 0;
      D.ekp[e] = i;
    }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_run += t; }
    __syncthreads();
  }
  if (tid == 0) { D.ecount[s] = s_run; D.estart[s] = s * D.cap; }
}

// K-glue 3: after the first PoseOptimization: drop outlier MapPoints (src/Tracking.cc:2402-2419), mark the
// keypoints and MapPoints that stay matched (SearchLocalPoints skips them, :2864-2869)
__global__ void __launch_bounds__(128) after_pose1_kernel(const TrackDev D) {
  const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.cap) return;
  const size_t o = (size_t)s * D.cap + i;
  if (i >= D.n[D.ips * s]) { D.blocked[o] = 0; return; }
  const int q = D.curMatch[o];
  uint8_t blk = 0;
  if (q >= 0) {
    const int e = D.kpEdge[o];
    if (D.eoutlier[(size_t)s * D.cap + e]) D.curMatch[o] = -1;
    else { blk = 1; D.mpTaken[(size_t)s * D.mcap + q] = 1; }
  }
  D.blocked[o] = blk;
}

// K-glue 4: Frame::isInFrustum's projection (src/Frame.cc:581-657) with the pose of PoseOptimization #1
__global__ void __launch_bounds__(128) project_map_kernel(const TrackDev D, float minX, float minY, float maxX, float maxY) {
  const int s = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= D.cap) return;
  const size_t o = (size_t)s * D.mcap + q;
  uint8_t fl = 0;
  if (q < D.n[D.ips * s] && (D.mpFlags[o] & 1) && !D.mpTaken[o]) {
    const float* T = D.T1 + 16 * s;
    const float X = D.xw[3 * o], Y = D.xw[3 * o + 1], Z = D.xw[3 * o + 2];
    const float xc = T[0] * X + T[1] * Y + T[2] * Z + T[3];
    const float yc = T[4] * X + T[5] * Y + T[6] * Z + T[7];
    const float zc = T[8] * X + T[9] * Y + T[10] * Z + T[11];
    if (zc > 0.0f) {
      const float invz = 1.0f / zc;
      const float u = D.fx * xc / zc + D.cx, v = D.fy * yc / zc + D.cy;
      if (u >= minX && u <= maxX && v >= minY && v <= maxY) {
        D.projX[o] = u;
        D.projY[o] = v;
        D.projXR[o] = u - D.bf * invz;
        D.level[o] = D.octave[o];      // PredictScale: the point is seen at its own distance
        D.viewCos[o] = 1.0f;
        fl = 3;
      }
    }
  }
  D.mapFlags[o] = fl;
}

// K-glue 4b (given map): Tracking::SearchLocalPoints (src/Tracking.cc:2848-2967) = Frame::isInFrustum(pMP, 0.5) for every local
// MapPoint that is not already matched, with the pose of PoseOptimization #1.  Same arithmetic as frustum_kernel
// (orbx_frame.cu, src/Frame.cc:571-650 + MapPoint::PredictScale src/MapPoint.cc:578-594); Ow = -Rcw^T tcw per stream.
__global__ void __launch_bounds__(128) frustum_map_kernel(const TrackDev D, float minX, float minY, float maxX, float maxY) {
  const int s = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= D.mcap) return;
  const size_t o = (size_t)s * D.mcap + q;
  uint8_t fl = 0;
  if (q < D.nMap[s] && (D.gMapFlags[o] & 1) && !D.mpTaken[o]) {
    const float* T = D.T1 + 16 * s;
    // Frame::UpdatePoseMatrices: mOw = -mRcw.t()*mtcw (src/Frame.cc:543) -- the transposed operand sends cv::gemm down its
    // general path: products and sum in double, rounded to float once
    const float Ow0 = (float)(-((double)T[0] * (double)T[3] + (double)T[4] * (double)T[7] + (double)T[8] * (double)T[11])),
                Ow1 = (float)(-((double)T[1] * (double)T[3] + (double)T[5] * (double)T[7] + (double)T[9] * (double)T[11])),
                Ow2 = (float)(-((double)T[2] * (double)T[3] + (double)T[6] * (double)T[7] + (double)T[10] * (double)T[11]));
    const float X = D.xw[3 * o], Y = D.xw[3 * o + 1], Z = D.xw[3 * o + 2];
    const float xc = T[0] * X + T[1] * Y + T[2] * Z + T[3];
    const float yc = T[4] * X + T[5] * Y + T[6] * Z + T[7];
    const float zc = T[8] * X + T[9] * Y + T[10] * Z + T[11];
    if (!(zc < 0.0f)) {
      const float invz = 1.0f / zc;
      const float u = D.fx * xc / zc + D.cx, v = D.fy * yc / zc + D.cy;
      if (!(u < minX || u > maxX) && !(v < minY || v > maxY)) {
        const float maxDistance = 1.2f * D.gMaxDist[o], minDistance = 0.8f * D.gMinDist[o];
        const float P0 = X - Ow0, P1 = Y - Ow1, P2 = Z - Ow2;
        const float dist = (float)sqrt((double)P0 * (double)P0 + (double)P1 * (double)P1 + (double)P2 * (double)P2);
        if (!(dist < minDistance || dist > maxDistance)) {
          const double dot = (double)P0 * (double)D.gNormal[3 * o] + (double)P1 * (double)D.gNormal[3 * o + 1] + (double)P2 * (double)D.gNormal[3 * o + 2];
          const float viewCos = (float)(dot / (double)dist);
          if (!(viewCos < 0.5f)) {
            const float ratio = D.gMaxDist[o] / dist;
            int lvl = (int)ceil(log((double)ratio) / (double)D.logScale);
            if (lvl < 0) lvl = 0; else if (lvl >= D.nlevels) lvl = D.nlevels - 1;
            D.projX[o] = u;
            D.projY[o] = v;
            D.projXR[o] = u - D.bf * invz;
            D.level[o] = lvl;
            D.viewCos[o] = viewCos;
            fl = (uint8_t)(1 | (D.gMapFlags[o] & 2));
          }
        }
      }
    }
  }
  D.mapFlags[o] = fl;
}

// K-glue 0 (chain mode): prior = dT * Tlast (4x4 row-major, fp32, fixed order), one thread per stream
__global__ void chain_prior_kernel(const float* dT, const float* Tlast, float* prior, float* T1, int S) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const float *A = dT + 16 * s, *B = Tlast + 16 * s;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A[4 * i], B[j]), __fmul_rn(A[4 * i + 1], B[4 + j])), __fmul_rn(A[4 * i + 2], B[8 + j])),
                                __fmul_rn(A[4 * i + 3], B[12 + j]));
      prior[16 * s + 4 * i + j] = v;
      T1[16 * s + 4 * i + j] = v;
    }
}

// K-glue 5: MapPoint of each keypoint after the local-map search, then per-stream stats
__global__ void __launch_bounds__(256) merge_matches_kernel(const TrackDev D) {
  const int s = blockIdx.x, tid = threadIdx.x;
  const int n = D.n[D.ips * s];
  const size_t base = (size_t)s * D.cap, mb = (size_t)s * D.mcap;
  for (int i = tid; i < n; i += 256) D.kpMp[base + i] = D.curMatch[base + i];
  __syncthreads();
  // replay  F.mvpMapPoints[bestIdx[q]] = pMP_q.  Every MapPoint of this harness has Observations() > 0 (checked by
  // orbx_tracker_set_map's callers), so a keypoint is claimed by at most one query and the scatter order is immaterial.
  const int nq = D.useMap ? D.nMap[s] : n;
  for (int q = tid; q < nq; q += 256) {
    const int b = D.bestIdx[mb + q];
    if (b >= 0) D.kpMp[base + b] = q;
  }
}

__global__ void __launch_bounds__(256) stats_kernel(const TrackDev D) {
  const int s = blockIdx.x, tid = threadIdx.x;
  const int n = D.n[D.ips * s];
  const size_t base = (size_t)s * D.cap;
  __shared__ int s_st;
  if (tid == 0) s_st = 0;
  __syncthreads();
  int c = 0;
  for (int i = tid; i < n; i += 256) c += D.uright[base + i] >= 0;
  if (c) atomicAdd(&s_st, c);
  __syncthreads();
  if (tid == 0) {
    int* st = D.stats + ORBX_TRACK_STATS * s;
    st[0] = n;
    st[1] = (D.ips == 2 ? D.n[2 * s + 1] : 0);
    st[2] = s_st;
    st[3] = D.nm1[s];
    st[4] = D.ninl[s];
    st[5] = D.nm2[s];
    st[6] = D.ninl[D.S + s];
    int it = 0;
    for (int k = 0; k < 4; ++k) it += D.iters[4 * s + k] + D.iters[4 * D.S + 4 * s + k];
    st[7] = it;
  }
}

// One set of per-step device state.  The tracker owns two: in overlap mode step t runs on slot t%2, its
// extraction + stereo stage on stream A and its matching + pose stage on stream B, so the latency-bound fp64
// pose optimisation of step t overlaps the throughput-bound extraction of step t+1.
struct TrackSlot {
  TrackDev D{};
  orbx_keypoint* d_kps = nullptr;
  uint8_t* d_desc = nullptr;
  int *d_n = nullptr, *d_mono = nullptr;
  FrameDev* d_frames = nullptr;
  StereoArgs* d_stereo = nullptr;
  uint16_t* d_stereoIdx = nullptr;   // [S][cap] row-bucketed right keypoint ids (stereo_rows_kernel)
  int* d_stereoRows = nullptr;       // [S][ORBX_STEREO_MAX_ROWS + 1]
  SbpFrameArgs* d_sbpf = nullptr;
  SbpMapArgs* d_sbpm = nullptr;
  double* d_scratch = nullptr;
  int* d_misc = nullptr;      // per-frame candidate totals / error flags
  cudaEvent_t evA = nullptr, evB = nullptr;
  bool usedB = false;
  // host copies of the per-stream argument blocks (pointer fields follow the map binding) + the binding they hold
  std::vector<SbpFrameArgs> hF;
  std::vector<SbpMapArgs> hM;
  unsigned long long mapVersion = ~0ull;
  float* dT = nullptr;        // [S][16] chain mode: the relative motion handed to the step
  // device copy of a host-side map (orbx_tracker_upload_map), one per slot: the map of step t is read until stage B
  // of step t has finished, while step t+1's map is already being uploaded into the other slot
  uint8_t* mapBuf = nullptr;
  cudaEvent_t evMap = nullptr;
  // device copy of host-side inertial inputs (orbx_tracker_upload_inertial), per slot like the map
  uint8_t* imuBuf = nullptr;
  cudaEvent_t evImu = nullptr;
  bool imuPending = false;    // stream B has to wait for evImu before it reads imuBuf
};

struct orbx_tracker {
  orbx_ctx* ctx = nullptr;
  orbx_ext* ext = nullptr;
  cudaStream_t stA = nullptr, stB = nullptr;   // stB == stA unless overlap mode is on
  cudaStream_t ownB = nullptr;
  int S = 0, cap = 0, nlevels = 0;
  int ips = 2;                 // images per stream (1: monocular tracker)
  orbx_camera cam{};
  float thFrame = 7.f, thMap = 1.f, nnMap = 0.8f;
  TrackSlot slot[2];
  int nslots = 1;
  unsigned long long stepCount = 0;
  std::vector<void*> allocs;
  float* d_scale = nullptr;
  uint8_t* d_imgs = nullptr;  // staging for the host-pointer entry point
  float *d_hposeIn = nullptr, *d_hposeOut = nullptr;
  int* d_hstats = nullptr;
  uint8_t* h_imgs = nullptr;
  float* h_pose = nullptr;
  int* h_stats = nullptr;
  int argW = -1, argH = -1, argStride = -1;
  const uint8_t* argImgs = nullptr;
  bool profiling = false, profiled = false;
  cudaEvent_t ev[ORBX_TRACK_STAGES + 1] = {};
  // asynchronous host-buffer pipeline (orbx_tracker_submit / orbx_tracker_collect)
  cudaStream_t stC = nullptr;                       // copy stream: H2D of step t+1 runs under the kernels of step t
  cudaEvent_t evH2D = nullptr, evStaged = nullptr;  // staging buffer filled / consumed (copied into the pyramid's level 0)
  bool stagedUsed = false;
  float *d_aPoseIn[2] = {}, *d_aPoseOut[2] = {}, *h_aPoseIn[2] = {}, *h_aPoseOut[2] = {};
  int *d_aStats[2] = {}, *h_aStats[2] = {};
  cudaEvent_t evDone[2] = {};
  int ringSlot[2] = {0, 0};
  unsigned long long submitted = 0, collected = 0;
  // given map + pose chaining
  int mcap = 0;
  float logScale = 1.f;
  bool haveMap = false;
  orbx_track_map map{};
  unsigned long long mapVersion = 0;
  bool chain = false;
  float* d_Tlast = nullptr;
  // CUDA graph of one step (orbx_tracker_set_graph): the ~45 launches of a step replayed as ONE graph launch when the
  // step's arguments repeat (single-frame latency path)
  bool graphMode = false;
  struct GraphKey {
    const void *imgs, *Ttrue, *Tprior, *Tout, *stats;
    int w, h, stride;
    unsigned long long mapVersion;
    bool chain;
    bool operator==(const GraphKey& o) const {
      return imgs == o.imgs && Ttrue == o.Ttrue && Tprior == o.Tprior && Tout == o.Tout && stats == o.stats && w == o.w && h == o.h &&
             stride == o.stride && mapVersion == o.mapVersion && chain == o.chain;
    }
  } graphKey{}, graphSeen{};
  cudaGraphExec_t graphExec = nullptr;
  size_t graphKernels = 0;
  unsigned long long graphLaunches = 0;
  // visual-inertial TrackLocalMap (src/Tracking.cc:2466-2490): mode 0 off, 1 LastKeyFrame, 2 LastFrame
  orbx_track_imu imu{};
  void* d_imuArgs = nullptr;                   // argument blocks of the inertial kernels
  double *d_imuState = nullptr, *d_imuH = nullptr;   // [S][21], [S][225] results (also readable as the next step's prior)
  // keyframe-rate work (LocalMapping's share: CreateNewMapPoints searches + LocalBundleAdjustment of every stream),
  // enqueued every kfPeriod-th step on its own lowest-priority stream
  orbx_tri_batch* kfTri = nullptr;
  orbx_lba_batch* kfLba = nullptr;
  int kfPeriod = 0;
  cudaStream_t stK = nullptr;
  unsigned long long kfRuns = 0;
};

template <typename T>
static T* talloc(orbx_tracker* t, size_t count) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) return nullptr;
  // cudaMemset runs on the legacy default stream, which the tracker's NON-BLOCKING streams do not wait for: without the
  // synchronisation a lazily allocated buffer (the map / inertial staging of the first upload) could be zeroed AFTER the
  // copies that fill it had landed -- seen as a rare wrong result on a cold box and under compute-sanitizer
  cudaMemset(p, 0, std::max<size_t>(count, 1) * sizeof(T));
  cudaStreamSynchronize(0);
  t->allocs.push_back(p);
  return (T*)p;
}

static bool slot_alloc(orbx_tracker* t, TrackSlot& K, const float* isg, float thFrame, float thMap, float nnMap) {
  const int S = t->S, cap = t->cap, mcap = t->mcap;
  const size_t SC = (size_t)S * cap, SM = (size_t)S * mcap;
  const orbx_camera* cam = &t->cam;
  TrackDev& D = K.D;
  D.S = S;
  D.cap = cap;
  D.ips = t->ips;
  D.mcap = mcap;
  D.useMap = 0;
  D.nMap = nullptr;
  D.gMapFlags = nullptr; D.gMaxDist = D.gMinDist = D.gNormal = nullptr;
  D.nlevels = t->nlevels;
  D.logScale = t->logScale;
  D.fx = cam->fx; D.fy = cam->fy; D.cx = cam->cx; D.cy = cam->cy; D.bf = cam->bf;
  for (int l = 0; l < t->nlevels; ++l) D.invSigma2[l] = isg[l];
  K.d_kps = talloc<orbx_keypoint>(t, 2 * SC);
  K.d_desc = talloc<uint8_t>(t, 2 * SC * 32);
  K.d_n = talloc<int>(t, 2 * S);
  K.d_mono = talloc<int>(t, 2 * S);
  D.kps = K.d_kps; D.desc = K.d_desc; D.n = K.d_n;
  D.uright = talloc<float>(t, SC); D.depth = talloc<float>(t, SC);
  D.mpFlags = talloc<uint8_t>(t, SM); D.lastFlags = talloc<uint8_t>(t, SM); D.xwOwn = talloc<float>(t, 3 * SM);
  D.xw = D.xwOwn;
  D.octave = talloc<int>(t, SM); D.angle = talloc<float>(t, SM);
  D.matchIdx = talloc<int>(t, SM); D.curMatch = talloc<int>(t, SC); D.nm1 = talloc<int>(t, S);
  D.kept = talloc<uint8_t>(t, SM);
  K.dT = talloc<float>(t, 16 * S);
  D.exw = talloc<float>(t, 3 * SC); D.eobs = talloc<float>(t, 3 * SC); D.eisg = talloc<float>(t, SC);
  D.ekp = talloc<int>(t, SC); D.ecount = talloc<int>(t, S); D.estart = talloc<int>(t, S); D.kpEdge = talloc<int>(t, SC);
  D.eoutlier = talloc<uint8_t>(t, SC);
  D.eclose = talloc<uint8_t>(t, SC);
  D.ninl = talloc<int>(t, 2 * S); D.iters = talloc<int>(t, 8 * S);
  D.blocked = talloc<uint8_t>(t, SC); D.mpTaken = talloc<uint8_t>(t, SM); D.mapFlags = talloc<uint8_t>(t, SM);
  D.projX = talloc<float>(t, SM); D.projY = talloc<float>(t, SM); D.projXR = talloc<float>(t, SM); D.viewCos = talloc<float>(t, SM);
  D.level = talloc<int>(t, SM); D.bestIdx = talloc<int>(t, SM); D.nm2 = talloc<int>(t, S); D.kpMp = talloc<int>(t, SC);
  D.Ttrue = talloc<float>(t, 16 * S); D.Tprior = talloc<float>(t, 16 * S); D.T1 = talloc<float>(t, 16 * S); D.T2 = talloc<float>(t, 16 * S);
  D.stats = talloc<int>(t, ORBX_TRACK_STATS * S);
  K.d_frames = talloc<FrameDev>(t, S);
  K.d_stereo = talloc<StereoArgs>(t, S);
  K.d_stereoIdx = talloc<uint16_t>(t, (size_t)S * cap);
  K.d_stereoRows = talloc<int>(t, (size_t)S * (ORBX_STEREO_MAX_ROWS + 1));
  K.d_sbpf = talloc<SbpFrameArgs>(t, S);
  K.d_sbpm = talloc<SbpMapArgs>(t, S);
  K.d_scratch = talloc<double>(t, 3 * SC);
  K.d_misc = talloc<int>(t, 8 * S);
  int* cellStart = talloc<int>(t, (size_t)S * (ORBX_NCELLS + 1));
  int* cellIdx = talloc<int>(t, SC);
  const int candCap = mcap * 96;
  uint32_t* cand = talloc<uint32_t>(t, (size_t)S * candCap);
  int* candOfs = talloc<int>(t, SM);
  int* candCnt = talloc<int>(t, SM);
  if (!cand || !candCnt || !D.stats || !K.d_scratch || !K.d_misc || !K.dT) return false;
  if (cudaEventCreateWithFlags(&K.evA, cudaEventDisableTiming) != cudaSuccess) return false;
  if (cudaEventCreateWithFlags(&K.evB, cudaEventDisableTiming) != cudaSuccess) return false;
  // static parts of the per-frame argument blocks
  std::vector<FrameDev> F(S);
  std::vector<SbpFrameArgs> AF(S);
  std::vector<SbpMapArgs> AM(S);
  for (int s = 0; s < S; ++s) {
    const size_t o = (size_t)s * cap, m = (size_t)s * mcap;
    F[s].n = 0;
    F[s].nDev = K.d_n + t->ips * s;
    F[s].kps = K.d_kps + (size_t)(t->ips * s) * cap;
    F[s].desc = K.d_desc + (size_t)(t->ips * s) * cap * 32;
    F[s].uright = D.uright + o;
    F[s].cellStart = cellStart + (size_t)s * (ORBX_NCELLS + 1);
    F[s].cellIdx = cellIdx + o;
    F[s].minX = F[s].minY = 0; F[s].maxX = F[s].maxY = 1; F[s].wInv = F[s].hInv = 1;   // set per image size
    SbpFrameArgs& a = AF[s];
    a.nq = 0; a.nqDev = K.d_n + t->ips * s; a.TcDev = D.Tprior + 16 * s;
    a.flags = D.lastFlags + m; a.xw = D.xwOwn + 3 * m; a.octave = D.octave + m; a.angle = D.angle + m;
    a.mpDesc = F[s].desc;
    a.fx = cam->fx; a.fy = cam->fy; a.cx = cam->cx; a.cy = cam->cy; a.bf = cam->bf;
    a.th = thFrame; a.mode = 0; a.checkOri = 1; a.scaleFactors = t->d_scale;
    a.candOfs = candOfs + m; a.candCnt = candCnt + m; a.cand = cand + (size_t)s * candCap; a.candCap = candCap;
    a.total = K.d_misc + 8 * s; a.err = K.d_misc + 8 * s + 1;
    a.curBlocked = nullptr; a.matchIdx = D.matchIdx + m; a.kept = D.kept + m; a.curMatch = D.curMatch + o; a.nmatches = D.nm1 + s;
    SbpMapArgs& g = AM[s];
    g.nq = 0; g.nqDev = K.d_n + t->ips * s;
    g.projX = D.projX + m; g.projY = D.projY + m; g.projXR = D.projXR + m; g.viewCos = D.viewCos + m; g.level = D.level + m;
    g.mpDesc = F[s].desc; g.flags = D.mapFlags + m; g.th = thMap; g.nnratio = nnMap; g.scaleFactors = t->d_scale;
    g.candOfs = candOfs + m; g.candCnt = candCnt + m; g.cand = cand + (size_t)s * candCap; g.candCap = candCap;
    g.total = K.d_misc + 8 * s + 2; g.err = K.d_misc + 8 * s + 3;
    g.kpBlocked = D.blocked + o; g.bestIdx = D.bestIdx + m; g.nmatches = D.nm2 + s;
  }
  K.hF = AF;
  K.hM = AM;
  K.mapVersion = ~0ull;
  cudaMemcpy(K.d_frames, F.data(), sizeof(FrameDev) * S, cudaMemcpyHostToDevice);
  cudaMemcpy(K.d_sbpf, AF.data(), sizeof(SbpFrameArgs) * S, cudaMemcpyHostToDevice);
  cudaMemcpy(K.d_sbpm, AM.data(), sizeof(SbpMapArgs) * S, cudaMemcpyHostToDevice);
  cudaStreamSynchronize(0);   // pageable H2D on the legacy stream may still be in flight; the tracker's streams are non-blocking
  return cudaGetLastError() == cudaSuccess;
}

extern "C" {

void orbx_tracker_destroy(orbx_tracker* t) {
  if (!t) return;
  cudaSetDevice(t->ctx->device);
  cudaStreamSynchronize(t->stA);
  if (t->graphExec) cudaGraphExecDestroy(t->graphExec);
  if (t->ownB) cudaStreamSynchronize(t->ownB);
  for (void* p : t->allocs) cudaFree(p);
  if (t->h_imgs) cudaFreeHost(t->h_imgs);
  if (t->h_pose) cudaFreeHost(t->h_pose);
  if (t->h_stats) cudaFreeHost(t->h_stats);
  for (int i = 0; i <= ORBX_TRACK_STAGES; ++i)
    if (t->ev[i]) cudaEventDestroy(t->ev[i]);
  for (int k = 0; k < 2; ++k) {
    if (t->slot[k].evA) cudaEventDestroy(t->slot[k].evA);
    if (t->slot[k].evB) cudaEventDestroy(t->slot[k].evB);
    if (t->slot[k].evMap) cudaEventDestroy(t->slot[k].evMap);
    if (t->slot[k].evImu) cudaEventDestroy(t->slot[k].evImu);
  }
  if (t->ownB) cudaStreamDestroy(t->ownB);
  if (t->stK) {
    cudaStreamSynchronize(t->stK);
    cudaStreamDestroy(t->stK);
  }
  if (t->stC) {
    cudaStreamSynchronize(t->stC);
    cudaStreamDestroy(t->stC);
    cudaEventDestroy(t->evH2D);
    cudaEventDestroy(t->evStaged);
    for (int k = 0; k < 2; ++k) {
      if (t->evDone[k]) cudaEventDestroy(t->evDone[k]);
      if (t->h_aPoseIn[k]) cudaFreeHost(t->h_aPoseIn[k]);
      if (t->h_aPoseOut[k]) cudaFreeHost(t->h_aPoseOut[k]);
      if (t->h_aStats[k]) cudaFreeHost(t->h_aStats[k]);
    }
  }
  delete t;
}

static orbx_tracker* tracker_create(orbx_ctx* ctx, orbx_ext* ext, int S, const orbx_camera* cam, float th_frame, float th_map,
                                    float nnratio_map, int ips) {
  if (!ctx || !ext || S < 1 || !cam || (ips == 2 && !(cam->b > 0))) {
    orbx_set_error("orbx_tracker_create: invalid argument");
    return nullptr;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
  orbx_tracker* t = new orbx_tracker();
  t->ips = ips;
  t->ctx = ctx;
  t->ext = ext;
  t->stA = t->stB = (cudaStream_t)orbx_extractor_stream(ext);
  t->S = S;
  t->cap = orbx_extractor_max_keypoints(ext);
  t->mcap = std::max(t->cap, ORBX_TRACK_MAP_CAP);
  t->nlevels = orbx_extractor_levels(ext);
  t->cam = *cam;
  t->thFrame = th_frame;
  t->thMap = th_map;
  t->nnMap = nnratio_map;
  float sc[ORBX_MAX_LEVELS], isg[ORBX_MAX_LEVELS];
  orbx_extractor_scale_tables(ext, sc, nullptr, nullptr, isg);
  // Frame::mfLogScaleFactor = log(mfScaleFactor) (src/Frame.cc:99; the scale factor is level 1 of the extractor's table)
  t->logScale = t->nlevels > 1 ? (float)std::log((double)sc[1]) : 1.0f;
  t->d_scale = talloc<float>(t, ORBX_MAX_LEVELS);
  if (!t->d_scale || cudaMemcpy(t->d_scale, sc, sizeof(float) * t->nlevels, cudaMemcpyHostToDevice) != cudaSuccess ||
      !slot_alloc(t, t->slot[0], isg, th_frame, th_map, nnratio_map)) {
    orbx_set_error("orbx_tracker_create: device allocation failed");
    orbx_tracker_destroy(t);
    return nullptr;
  }
  return t;
}

orbx_tracker* orbx_tracker_create(orbx_ctx* ctx, orbx_ext* ext, int S, const orbx_camera* cam, float th_frame, float th_map,
                                  float nnratio_map) {
  return tracker_create(ctx, ext, S, cam, th_frame, th_map, nnratio_map, 2);
}

// Monocular tracker (BASELINE config 1; Frame::Frame(mono) src/Frame.cc:308-349, Tracking::TrackWithMotionModel with bMono,
// th = 15, src/Tracking.cc:2364-2378; monocular edges of PoseOptimization src/Optimizer.cc:961-1010): ONE image per stream,
// no ComputeStereoMatches (mvuRight = mvDepth = -1), the map must be given (orbx_tracker_set_map / _upload_map).
orbx_tracker* orbx_tracker_create_mono(orbx_ctx* ctx, orbx_ext* ext, int S, const orbx_camera* cam, float th_frame, float th_map,
                                       float nnratio_map) {
  return tracker_create(ctx, ext, S, cam, th_frame, th_map, nnratio_map, 1);
}
int orbx_tracker_images_per_stream(const orbx_tracker* t) { return t ? t->ips : ORBX_EINVAL; }

int orbx_tracker_set_overlap(orbx_tracker* t, int enable) {
  if (!t) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  ORBX_CUDA(cudaStreamSynchronize(t->stA));
  if (t->ownB) ORBX_CUDA(cudaStreamSynchronize(t->ownB));
  if (enable) {
    if (!t->ownB) {
      // stage B is light and latency-bound: give its stream the highest priority so its thread blocks take any
      // SM slot that frees up while the (throughput-bound) extraction grids of the next step drain
      int lo = 0, hi = 0;
      ORBX_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      int prio = hi;
      if (const char* pe = getenv("ORBX_STAGEB_PRIORITY")) {   // A/B switch for the scheduling study in DESIGN.md §7
        if (pe[0] == 'l') prio = lo;
        else if (pe[0] == 's') prio = 0;
      }
      ORBX_CUDA(cudaStreamCreateWithPriority(&t->ownB, cudaStreamNonBlocking, prio));
    }
    if (t->nslots < 2) {
      float sc[ORBX_MAX_LEVELS], isg[ORBX_MAX_LEVELS];
      orbx_extractor_scale_tables(t->ext, sc, nullptr, nullptr, isg);
      if (!slot_alloc(t, t->slot[1], isg, t->thFrame, t->thMap, t->nnMap)) {
        orbx_set_error("orbx_tracker_set_overlap: device allocation failed");
        return ORBX_ECUDA;
      }
      t->nslots = 2;
      t->argW = -1;   // rebind both slots' geometry
    }
    t->stB = t->ownB;
  } else {
    t->stB = t->stA;
  }
  t->slot[0].usedB = t->slot[1].usedB = false;
  return ORBX_OK;
}

void* orbx_tracker_result_stream(orbx_tracker* t) { return t ? (void*)t->stB : nullptr; }

int orbx_tracker_synchronize(orbx_tracker* t) {
  if (!t) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  ORBX_CUDA(cudaStreamSynchronize(t->stA));
  if (t->stB != t->stA) ORBX_CUDA(cudaStreamSynchronize(t->stB));
  if (t->stK) ORBX_CUDA(cudaStreamSynchronize(t->stK));
  return ORBX_OK;
}

// Keyframe-rate work.  In the reference LocalMapping runs in its own thread beside Tracking (src/LocalMapping.cc:68-200):
// a new keyframe triggers CreateNewMapPoints (one SearchForTriangulation per covisible neighbour, :501-628) and
// Optimizer::LocalBundleAdjustment (:201) while frames keep being tracked.  Here the two prepared many-stream plans are
// enqueued every `period`-th step on a third stream of the lowest priority, so they fill the SMs the per-frame chain
// leaves idle; a new round is only enqueued behind the previous one (stream order).
int orbx_tracker_set_keyframe_work(orbx_tracker* t, orbx_tri_batch* tri, orbx_lba_batch* lba, int period) {
  if (!t || period < 0 || ((tri || lba) && period < 1)) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  if (t->stK) ORBX_CUDA(cudaStreamSynchronize(t->stK));
  if ((tri || lba) && !t->stK) {
    int lo = 0, hi = 0;
    ORBX_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    ORBX_CUDA(cudaStreamCreateWithPriority(&t->stK, cudaStreamNonBlocking, lo));
  }
  t->kfTri = tri;
  t->kfLba = lba;
  t->kfPeriod = (tri || lba) ? period : 0;
  t->kfRuns = 0;
  return ORBX_OK;
}

// ---- visual-inertial TrackLocalMap ----
static bool imu_complete(const orbx_track_imu* u) {
  if (u->mode != 1 && u->mode != 2) return false;
  if (!u->Tcb || !u->Tbc || !u->velocity || !u->bias || !u->ref_state || !u->preint || !u->info_inertial || !u->info_gyro || !u->info_acc)
    return false;
  if (u->mode == 2 && (!u->preint_jac || !u->preint_bias || !u->prior_state || !u->prior_H)) return false;
  return true;
}
static int imu_ensure_buffers(orbx_tracker* t) {
  if (t->d_imuArgs) return ORBX_OK;
  t->d_imuArgs = talloc<uint8_t>(t, orbx_inertial_args_bytes(t->S));
  t->d_imuState = talloc<double>(t, 21 * (size_t)t->S);
  t->d_imuH = talloc<double>(t, 225 * (size_t)t->S);
  if (!t->d_imuArgs || !t->d_imuState || !t->d_imuH) return ORBX_ECUDA;
  ORBX_CUDA(cudaMemset(t->d_imuState, 0, sizeof(double) * 21 * t->S));
  ORBX_CUDA(cudaMemset(t->d_imuH, 0, sizeof(double) * 225 * t->S));
  ORBX_CUDA(cudaStreamSynchronize(0));
  return ORBX_OK;
}

// Bind DEVICE-resident inertial inputs for the following steps (NULL or mode 0: back to the visual PoseOptimization).
// The arrays are read when the step's stage B runs; they must stay valid and unchanged until then.
int orbx_tracker_set_inertial(orbx_tracker* t, const orbx_track_imu* imu) {
  if (!t) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  if (!imu || imu->mode == 0) {
    t->imu = orbx_track_imu{};
    return ORBX_OK;
  }
  if (!imu_complete(imu)) {
    orbx_set_error("orbx_tracker_set_inertial: mode must be 1 or 2 and every array of that mode non-null");
    return ORBX_EINVAL;
  }
  int rc = imu_ensure_buffers(t);
  if (rc != ORBX_OK) return rc;
  t->imu = *imu;
  return ORBX_OK;
}

// HOST-side inertial inputs -> the next step's slot (asynchronous, on the copy stream when the submit/collect pipeline is
// in use), then bound like orbx_tracker_set_inertial.  prior_state / prior_H may be NULL in mode 2: the tracker then uses
// the state and the marginalised Hessian the PREVIOUS step left on the device (the reference's pFp->mpcpi chain,
// src/Optimizer.cc:8594-8599).
int orbx_tracker_upload_inertial(orbx_tracker* t, const orbx_track_imu* h) {
  if (!t || !h) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  int rc = imu_ensure_buffers(t);
  if (rc != ORBX_OK) return rc;
  orbx_track_imu probe = *h;
  if (h->mode == 2 && !h->prior_state && !h->prior_H) { probe.prior_state = t->d_imuState; probe.prior_H = t->d_imuH; }
  if (h->mode == 2 && !h->ref_state) probe.ref_state = t->d_imuState;
  if (!imu_complete(&probe)) {
    orbx_set_error("orbx_tracker_upload_inertial: mode must be 1 or 2 and every array of that mode non-null");
    return ORBX_EINVAL;
  }
  const bool overlap = t->stB != t->stA;
  TrackSlot& K = t->slot[overlap ? (int)(t->stepCount & 1) : 0];
  const size_t S = (size_t)t->S;
  // layout (bytes): Tcb, Tbc | vel | bias | ref | preint | jac | pbias | infoI | infoG | infoA | priorState | priorH
  const size_t oT = 0, oV = 128, oB = align_up(oV + 12 * S, 256), oR = align_up(oB + 24 * S, 256), oP = oR + 8 * 21 * S,
               oJ = oP + 8 * 16 * S, oPb = oJ + 8 * 45 * S, oI = oPb + 8 * 6 * S, oG = oI + 8 * 81 * S, oA = oG + 8 * 9 * S,
               oPs = oA + 8 * 9 * S, oPh = oPs + 8 * 21 * S, total = oPh + 8 * 225 * S;
  if (!K.imuBuf) {
    K.imuBuf = talloc<uint8_t>(t, total);
    if (!K.imuBuf || cudaEventCreateWithFlags(&K.evImu, cudaEventDisableTiming) != cudaSuccess) return ORBX_ECUDA;
  }
  cudaStream_t sc = t->stC ? t->stC : t->stA;
  if (overlap && K.usedB) ORBX_CUDA(cudaStreamWaitEvent(sc, K.evB, 0));   // the slot's previous inputs are no longer read
  uint8_t* b = K.imuBuf;
  auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
    return src ? cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyHostToDevice, sc) : cudaSuccess;
  };
  ORBX_CUDA(up(oT, h->Tcb, 64));
  ORBX_CUDA(up(oT + 64, h->Tbc, 64));
  ORBX_CUDA(up(oV, h->velocity, 12 * S));
  ORBX_CUDA(up(oB, h->bias, 24 * S));
  ORBX_CUDA(up(oR, h->ref_state, 8 * 21 * S));
  ORBX_CUDA(up(oP, h->preint, 8 * 16 * S));
  ORBX_CUDA(up(oI, h->info_inertial, 8 * 81 * S));
  ORBX_CUDA(up(oG, h->info_gyro, 8 * 9 * S));
  ORBX_CUDA(up(oA, h->info_acc, 8 * 9 * S));
  if (h->mode == 2) {
    ORBX_CUDA(up(oJ, h->preint_jac, 8 * 45 * S));
    ORBX_CUDA(up(oPb, h->preint_bias, 8 * 6 * S));
    ORBX_CUDA(up(oPs, h->prior_state, 8 * 21 * S));
    ORBX_CUDA(up(oPh, h->prior_H, 8 * 225 * S));
  }
  ORBX_CUDA(cudaEventRecord(K.evImu, sc));
  K.imuPending = true;
  orbx_track_imu d{};
  d.mode = h->mode; d.rec_init = h->rec_init;
  d.Tcb = (const float*)(b + oT); d.Tbc = (const float*)(b + oT + 64);
  d.velocity = (const float*)(b + oV); d.bias = (const float*)(b + oB);
  d.ref_state = (h->ref_state || h->mode != 2) ? (const double*)(b + oR) : t->d_imuState; d.preint = (const double*)(b + oP); d.preint_jac = (const double*)(b + oJ);
  d.preint_bias = (const double*)(b + oPb); d.info_inertial = (const double*)(b + oI); d.info_gyro = (const double*)(b + oG);
  d.info_acc = (const double*)(b + oA);
  d.prior_state = h->prior_state ? (const double*)(b + oPs) : t->d_imuState;
  d.prior_H = h->prior_H ? (const double*)(b + oPh) : t->d_imuH;
  t->imu = d;
  return ORBX_OK;
}

// Results of the last inertial step: the optimised body states [S][21] (Rwb, twb, v, bg, ba) and the 15x15 Hessians
// [S][225] for the next ConstraintPoseImu; either pointer may be NULL.  Synchronises the tracker's streams.
int orbx_tracker_inertial_result(orbx_tracker* t, double* state, double* H15) {
  if (!t || !t->d_imuState) return ORBX_EINVAL;
  int rc = orbx_tracker_synchronize(t);
  if (rc != ORBX_OK) return rc;
  if (state) ORBX_CUDA(cudaMemcpy(state, t->d_imuState, sizeof(double) * 21 * t->S, cudaMemcpyDeviceToHost));
  if (H15) ORBX_CUDA(cudaMemcpy(H15, t->d_imuH, sizeof(double) * 225 * t->S, cudaMemcpyDeviceToHost));
  return ORBX_OK;
}
// Device addresses of those two arrays (valid for the tracker's lifetime once an inertial mode was set): bind them as
// prior_state / prior_H of the next LastFrame step to chain the prior on the device.
const double* orbx_tracker_inertial_state_dev(orbx_tracker* t) { return (t && imu_ensure_buffers(t) == ORBX_OK) ? t->d_imuState : nullptr; }
const double* orbx_tracker_inertial_hessian_dev(orbx_tracker* t) { return (t && imu_ensure_buffers(t) == ORBX_OK) ? t->d_imuH : nullptr; }

int orbx_tracker_map_capacity(const orbx_tracker* t) { return t ? t->mcap : ORBX_EINVAL; }

int orbx_tracker_set_map(orbx_tracker* t, const orbx_track_map* map) {
  if (!t) return ORBX_EINVAL;
  if (map) {
    if (map->m_cap != t->mcap || !map->n_map || !map->xw || !map->desc || !map->last_flags || !map->last_octave || !map->last_angle ||
        !map->map_flags || !map->max_dist || !map->min_dist || !map->normal) {
      orbx_set_error("orbx_tracker_set_map: incomplete map, or m_cap != orbx_tracker_map_capacity() = %d", t->mcap);
      return ORBX_EINVAL;
    }
    // the same arrays as the binding in place (e.g. orbx_tracker_upload_map refilling its slot buffer): the argument blocks
    // already point there, nothing to rebind
    const orbx_track_map& c = t->map;
    if (t->haveMap && c.m_cap == map->m_cap && c.n_map == map->n_map && c.xw == map->xw && c.desc == map->desc &&
        c.last_flags == map->last_flags && c.last_octave == map->last_octave && c.last_angle == map->last_angle &&
        c.map_flags == map->map_flags && c.max_dist == map->max_dist && c.min_dist == map->min_dist && c.normal == map->normal &&
        c.log_scale_factor == map->log_scale_factor)
      return ORBX_OK;
    t->map = *map;
    t->haveMap = true;
  } else {
    if (!t->haveMap) return ORBX_OK;
    t->haveMap = false;
  }
  t->mapVersion++;
  return ORBX_OK;
}

// Host-side map -> the next step's slot (page-locked caller memory is DMA'd directly, on the copy stream when the
// asynchronous pipeline is in use so that it overlaps the previous step's kernels), then bound like orbx_tracker_set_map.
int orbx_tracker_upload_map(orbx_tracker* t, const orbx_track_map* hm) {
  if (!t || !hm) return ORBX_EINVAL;
  if (hm->m_cap != t->mcap || !hm->n_map || !hm->xw || !hm->desc || !hm->last_flags || !hm->last_octave || !hm->last_angle ||
      !hm->map_flags || !hm->max_dist || !hm->min_dist || !hm->normal) {
    orbx_set_error("orbx_tracker_upload_map: incomplete map, or m_cap != orbx_tracker_map_capacity() = %d", t->mcap);
    return ORBX_EINVAL;
  }
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  const bool overlap = t->stB != t->stA;
  TrackSlot& K = t->slot[overlap ? (int)(t->stepCount & 1) : 0];
  const size_t S = (size_t)t->S, SM = S * t->mcap;
  // layout of the slot's map buffer
  const size_t oN = 0, oXw = align_up(oN + 4 * S, 256), oDesc = align_up(oXw + 12 * SM, 256), oLf = align_up(oDesc + 32 * SM, 256),
               oLo = align_up(oLf + SM, 256), oLa = align_up(oLo + 4 * SM, 256), oMf = align_up(oLa + 4 * SM, 256),
               oMx = align_up(oMf + SM, 256), oMn = align_up(oMx + 4 * SM, 256), oNr = align_up(oMn + 4 * SM, 256),
               total = align_up(oNr + 12 * SM, 256);
  if (!K.mapBuf) {
    K.mapBuf = talloc<uint8_t>(t, total);
    if (!K.mapBuf || cudaEventCreateWithFlags(&K.evMap, cudaEventDisableTiming) != cudaSuccess) return ORBX_ECUDA;
  }
  cudaStream_t sc = t->stC ? t->stC : t->stA;
  if (overlap && K.usedB) ORBX_CUDA(cudaStreamWaitEvent(sc, K.evB, 0));   // the slot's previous map is no longer read
  uint8_t* b = K.mapBuf;
  ORBX_CUDA(cudaMemcpyAsync(b + oN, hm->n_map, 4 * S, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaMemcpyAsync(b + oXw, hm->xw, 12 * SM, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaMemcpyAsync(b + oDesc, hm->desc, 32 * SM, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaMemcpyAsync(b + oLf, hm->last_flags, SM, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaMemcpyAsync(b + oLo, hm->last_octave, 4 * SM, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaMemcpyAsync(b + oLa, hm->last_angle, 4 * SM, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaMemcpyAsync(b + oMf, hm->map_flags, SM, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaMemcpyAsync(b + oMx, hm->max_dist, 4 * SM, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaMemcpyAsync(b + oMn, hm->min_dist, 4 * SM, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaMemcpyAsync(b + oNr, hm->normal, 12 * SM, cudaMemcpyHostToDevice, sc));
  if (sc != t->stA) {
    ORBX_CUDA(cudaEventRecord(K.evMap, sc));
    ORBX_CUDA(cudaStreamWaitEvent(t->stA, K.evMap, 0));
  }
  orbx_track_map dm;
  dm.m_cap = t->mcap;
  dm.n_map = (const int32_t*)(b + oN); dm.xw = (const float*)(b + oXw); dm.desc = b + oDesc; dm.last_flags = b + oLf;
  dm.last_octave = (const int32_t*)(b + oLo); dm.last_angle = (const float*)(b + oLa); dm.map_flags = b + oMf;
  dm.max_dist = (const float*)(b + oMx); dm.min_dist = (const float*)(b + oMn); dm.normal = (const float*)(b + oNr);
  dm.log_scale_factor = hm->log_scale_factor;
  return orbx_tracker_set_map(t, &dm);
}
size_t orbx_tracker_map_bytes(const orbx_tracker* t) {
  if (!t) return 0;
  const size_t S = (size_t)t->S, SM = S * t->mcap;
  return 4 * S + SM * (12 + 32 + 1 + 4 + 4 + 1 + 4 + 4 + 12);
}

int orbx_tracker_set_chain(orbx_tracker* t, int enable, const float* d_Tcw_init) {
  if (!t || (enable && !d_Tcw_init)) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  int rc = orbx_tracker_synchronize(t);
  if (rc != ORBX_OK) return rc;
  if (enable) {
    if (!t->d_Tlast) t->d_Tlast = talloc<float>(t, 16 * (size_t)t->S);
    if (!t->d_Tlast) return ORBX_ECUDA;
    ORBX_CUDA(cudaMemcpy(t->d_Tlast, d_Tcw_init, sizeof(float) * 16 * t->S, cudaMemcpyDefault));   // host or device pointer
    ORBX_CUDA(cudaStreamSynchronize(0));
  }
  t->chain = enable != 0;
  return ORBX_OK;
}

void* orbx_tracker_keyframe_stream(orbx_tracker* t) { return t ? (void*)t->stK : nullptr; }
long long orbx_tracker_keyframe_runs(const orbx_tracker* t) { return t ? (long long)t->kfRuns : -1; }

// image-size dependent argument blocks (stereo pyramids, grid bounds): rebuilt when (w,h) or the input binding changes
static int tracker_bind_geometry(orbx_tracker* t, int w, int h) {
  const int S = t->S;
  ORBX_CUDA(cudaStreamSynchronize(t->stA));
  if (t->stB != t->stA) ORBX_CUDA(cudaStreamSynchronize(t->stB));
  for (int k = 0; k < t->nslots; ++k) {
    TrackSlot& K = t->slot[k];
    std::vector<StereoArgs> A(S);
    std::vector<FrameDev> F(S);
    ORBX_CUDA(cudaMemcpy(F.data(), K.d_frames, sizeof(FrameDev) * S, cudaMemcpyDeviceToHost));
    for (int s = 0; s < S; ++s) {
      F[s].minX = 0.f; F[s].minY = 0.f; F[s].maxX = (float)w; F[s].maxY = (float)h;   // undistorted image bounds (src/Frame.cc:147-152)
      F[s].wInv = (float)ORBX_GRID_COLS / (F[s].maxX - F[s].minX);
      F[s].hInv = (float)ORBX_GRID_ROWS / (F[s].maxY - F[s].minY);
      if (t->ips != 2) continue;
      StereoArgs& a = A[s];
      int nl = 0, nl2 = 0, w2[ORBX_MAX_LEVELS], h2[ORBX_MAX_LEVELS];
      float sc2[ORBX_MAX_LEVELS], isc2[ORBX_MAX_LEVELS];
      cudaStream_t st1, st2;
      int rc = orbx_ext_pyramid_view(t->ext, 2 * s, &nl, a.pyrL, a.lw, a.lh, a.pitchL, a.scale, a.invScale, &st1);
      if (rc != ORBX_OK) return rc;
      rc = orbx_ext_pyramid_view(t->ext, 2 * s + 1, &nl2, a.pyrR, w2, h2, a.pitchR, sc2, isc2, &st2);
      if (rc != ORBX_OK) return rc;
      a.nlevels = nl;
      a.nL = a.nR = 0;
      a.nLDev = K.d_n + 2 * s;
      a.nRDev = K.d_n + 2 * s + 1;
      a.kpL = K.d_kps + (size_t)(2 * s) * t->cap;
      a.kpR = K.d_kps + (size_t)(2 * s + 1) * t->cap;
      a.descL = K.d_desc + (size_t)(2 * s) * t->cap * 32;
      a.descR = K.d_desc + (size_t)(2 * s + 1) * t->cap * 32;
      a.bf = t->cam.bf;
      a.b = t->cam.b;
      a.uright = K.D.uright + (size_t)s * t->cap;
      a.depth = K.D.depth + (size_t)s * t->cap;
      a.sad = K.D.kpEdge + (size_t)s * t->cap;   // scratch reuse: kpEdge is rewritten later in the step
      a.sortIdx = K.d_stereoIdx + (size_t)s * t->cap;
      a.rowStart = K.d_stereoRows + (size_t)s * (ORBX_STEREO_MAX_ROWS + 1);
      F[s].minX = 0.f; F[s].minY = 0.f; F[s].maxX = (float)w; F[s].maxY = (float)h;   // rectified stereo: image bounds (src/Frame.cc:147-152)
      F[s].wInv = (float)ORBX_GRID_COLS / (F[s].maxX - F[s].minX);
      F[s].hInv = (float)ORBX_GRID_ROWS / (F[s].maxY - F[s].minY);
    }
    ORBX_CUDA(cudaMemcpy(K.d_stereo, A.data(), sizeof(StereoArgs) * S, cudaMemcpyHostToDevice));
    ORBX_CUDA(cudaMemcpy(K.d_frames, F.data(), sizeof(FrameDev) * S, cudaMemcpyHostToDevice));
  }
  ORBX_CUDA(cudaStreamSynchronize(0));   // see slot_alloc
  t->argW = w;
  t->argH = h;
  return ORBX_OK;
}

static int tracker_step_body(orbx_tracker* t, const uint8_t* d_imgs, int w, int h, int stride, const float* d_Tcw_true,
                             const float* d_Tcw_prior, float* d_Tcw_out, int32_t* d_stats) {
  cudaStream_t sa = t->stA, sb = t->stB;
  const bool overlap = sb != sa;
  const int S = t->S, cap = t->cap;
  TrackSlot& K = t->slot[overlap ? (int)(t->stepCount & 1) : 0];
  TrackDev& D = K.D;
  cudaEvent_t* ev = t->profiling ? t->ev : nullptr;
#define TRK_EV(i, st) do { if (ev) ORBX_CUDA(cudaEventRecord(ev[i], st)); } while (0)
  // the slot's buffers are free once stage B of the step that last used them has finished
  if (overlap && K.usedB) ORBX_CUDA(cudaStreamWaitEvent(sa, K.evB, 0));
  if (K.mapVersion != t->mapVersion) {
    // (re)bind the map-side pointers of this slot's argument blocks: the caller's flat map or the tracker's own self-map
    const bool gm = t->haveMap;
    const orbx_track_map& M = t->map;
    D.useMap = gm ? 1 : 0;
    D.nMap = gm ? M.n_map : nullptr;
    D.xw = gm ? M.xw : D.xwOwn;
    D.gMapFlags = gm ? M.map_flags : nullptr;
    D.gMaxDist = gm ? M.max_dist : nullptr;
    D.gMinDist = gm ? M.min_dist : nullptr;
    D.gNormal = gm ? M.normal : nullptr;
    if (gm && M.log_scale_factor > 0) D.logScale = M.log_scale_factor;
    for (int s = 0; s < S; ++s) {
      const size_t m = (size_t)s * t->mcap;
      SbpFrameArgs& a = K.hF[s];
      SbpMapArgs& g = K.hM[s];
      const uint8_t* ownDesc = K.d_desc + (size_t)(t->ips * s) * cap * 32;
      a.nqDev = gm ? M.n_map + s : K.d_n + t->ips * s;
      a.flags = gm ? M.last_flags + m : D.lastFlags + m;
      a.xw = D.xw + 3 * m;
      a.octave = gm ? M.last_octave + m : D.octave + m;
      a.angle = gm ? M.last_angle + m : D.angle + m;
      a.mpDesc = gm ? M.desc + 32 * m : ownDesc;
      g.nqDev = a.nqDev;
      g.mpDesc = a.mpDesc;
    }
    ORBX_CUDA(cudaMemcpyAsync(K.d_sbpf, K.hF.data(), sizeof(SbpFrameArgs) * S, cudaMemcpyHostToDevice, sa));
    ORBX_CUDA(cudaMemcpyAsync(K.d_sbpm, K.hM.data(), sizeof(SbpMapArgs) * S, cudaMemcpyHostToDevice, sa));
    K.mapVersion = t->mapVersion;
  }
  TRK_EV(0, sa);
  // 1. Frame::Frame(stereo): ORB extraction of the 2*S images (src/Frame.cc:111-114)
  if (t->ips == 1 && !t->haveMap) {
    orbx_set_error("orbx_tracker: the monocular tracker has no self-map harness (no depth): set a map first");
    return ORBX_EINVAL;
  }
  int rc = orbx_extract_batch_device(t->ext, t->ips * S, d_imgs, w, h, stride, 0, 0, K.d_kps, K.d_desc, cap, K.d_n, K.d_mono);
  if (rc != ORBX_OK) return rc;
  if (w != t->argW || h != t->argH || stride != t->argStride || d_imgs != t->argImgs) {   // level 0 aliases the input
    rc = tracker_bind_geometry(t, w, h);
    if (rc != ORBX_OK) return rc;
    t->argStride = stride;
    t->argImgs = d_imgs;
  }
  TRK_EV(1, sa);
  ORBX_CUDA(cudaMemcpyAsync(D.Ttrue, d_Tcw_true, sizeof(float) * 16 * S, cudaMemcpyDeviceToDevice, sa));
  if (t->chain) {
    ORBX_CUDA(cudaMemcpyAsync(K.dT, d_Tcw_prior, sizeof(float) * 16 * S, cudaMemcpyDeviceToDevice, sa));   // relative motion
  } else {
    ORBX_CUDA(cudaMemcpyAsync(D.Tprior, d_Tcw_prior, sizeof(float) * 16 * S, cudaMemcpyDeviceToDevice, sa));
    ORBX_CUDA(cudaMemcpyAsync(D.T1, d_Tcw_prior, sizeof(float) * 16 * S, cudaMemcpyDeviceToDevice, sa));
  }
  ORBX_CUDA(cudaMemsetAsync(K.d_misc, 0, sizeof(int) * 8 * S, sa));
  ORBX_CUDA(cudaMemsetAsync(D.mpTaken, 0, (size_t)S * t->mcap, sa));
  // 2. ComputeStereoMatches (src/Frame.cc:132)
  if (t->ips == 2) {
    rc = orbx_launch_stereo_batch(t->ctx, sa, K.d_stereo, S, cap);
    if (rc != ORBX_OK) return rc;
  } else {
    mono_fill_kernel<<<dim3(div_up(cap, 128), S), 128, 0, sa>>>(D);
    ORBX_LAUNCH(t->ctx);
  }
  if (overlap) {
    ORBX_CUDA(cudaEventRecord(K.evA, sa));
    ORBX_CUDA(cudaStreamWaitEvent(sb, K.evA, 0));
  }
  TRK_EV(2, sb);
  // 3. synthetic map + AssignFeaturesToGrid + SearchByProjection(Cur, Last, th) (src/Tracking.cc:2370)
  const dim3 gk(div_up(cap, 128), S), gm(div_up(t->mcap, 128), S);
  const int maxQ = D.useMap ? t->mcap : cap;
  if (t->chain) {   // motion model: prior = dT * (pose the previous step produced); stage B of that step precedes us on sb
    chain_prior_kernel<<<div_up(S, 128), 128, 0, sb>>>(K.dT, t->d_Tlast, D.Tprior, D.T1, S);
    ORBX_LAUNCH(t->ctx);
  }
  if (!D.useMap) {
    backproject_kernel<<<gk, 128, 0, sb>>>(D);
    ORBX_LAUNCH(t->ctx);
  }
  rc = orbx_launch_grid_build(t->ctx, sb, K.d_frames, S);
  if (rc != ORBX_OK) return rc;
  rc = orbx_launch_sbp_frame_batch(t->ctx, sb, K.d_frames, K.d_sbpf, S, maxQ, cap);
  if (rc != ORBX_OK) return rc;
  TRK_EV(3, sb);
  // 4. PoseOptimization (src/Tracking.cc:2395)
  gather_edges_kernel<<<S, 256, 0, sb>>>(D, 0);
  ORBX_LAUNCH(t->ctx);
  rc = orbx_launch_pose_opt_slices(t->ctx, sb, S, D.estart, D.ecount, D.exw, D.eobs, D.eisg, &t->cam, D.T1, D.eoutlier, D.ninl,
                                   D.iters, K.d_scratch);
  if (rc != ORBX_OK) return rc;
  TRK_EV(4, sb);
  // 5. TrackLocalMap: SearchLocalPoints + SearchByProjection(F, local points, th) (src/Tracking.cc:2449,:2964)
  after_pose1_kernel<<<gk, 128, 0, sb>>>(D);
  ORBX_LAUNCH(t->ctx);
  if (D.useMap) frustum_map_kernel<<<gm, 128, 0, sb>>>(D, 0.f, 0.f, (float)w, (float)h);
  else project_map_kernel<<<gk, 128, 0, sb>>>(D, 0.f, 0.f, (float)w, (float)h);
  ORBX_LAUNCH(t->ctx);
  rc = orbx_launch_sbp_map_batch(t->ctx, sb, K.d_frames, K.d_sbpm, S, maxQ, cap);
  if (rc != ORBX_OK) return rc;
  TRK_EV(5, sb);
  // 6. PoseOptimization (src/Tracking.cc:2468), starting from the pose of step 4
  merge_matches_kernel<<<S, 256, 0, sb>>>(D);
  ORBX_LAUNCH(t->ctx);
  gather_edges_kernel<<<S, 256, 0, sb>>>(D, 1);
  ORBX_LAUNCH(t->ctx);
  // inliers / iterations of the second optimisation land in the second halves of ninl / iters
  if (t->imu.mode) {
    // visual-inertial TrackLocalMap (src/Tracking.cc:2466-2490): PoseInertialOptimizationLastKeyFrame / LastFrame on the
    // same edges; the frame's velocity, bias, the reference state and the pre-integration come from orbx_tracker_set_inertial
    if (K.imuPending) {
      ORBX_CUDA(cudaStreamWaitEvent(sb, K.evImu, 0));
      K.imuPending = false;
    }
    const orbx_track_imu& U = t->imu;
    OrbxInertialSlices I{};
    I.mode = U.mode; I.S = S; I.recInit = U.rec_init;
    I.fx = t->cam.fx; I.fy = t->cam.fy; I.cx = t->cam.cx; I.cy = t->cam.cy; I.bf = t->cam.bf;
    I.estart = D.estart; I.ecount = D.ecount; I.exw = D.exw; I.eobs = D.eobs; I.eisg = D.eisg; I.eclose = D.eclose;
    I.eoutlier = D.eoutlier; I.err = K.d_scratch;
    I.T1 = D.T1; I.Tcb = U.Tcb; I.Tbc = U.Tbc; I.vel = U.velocity; I.bias = U.bias;
    I.ref = U.ref_state; I.preint = U.preint; I.preintJac = U.preint_jac; I.preintBias = U.preint_bias;
    I.infoI = U.info_inertial; I.infoG = U.info_gyro; I.infoA = U.info_acc; I.priorState = U.prior_state; I.priorH = U.prior_H;
    I.stateOut = t->d_imuState; I.H15 = t->d_imuH; I.nRet = D.ninl + S; I.iters = D.iters + 4 * S; I.T2 = D.T2;
    rc = orbx_launch_pose_inertial_slices(t->ctx, sb, I, t->d_imuArgs);
  } else {
    ORBX_CUDA(cudaMemcpyAsync(D.T2, D.T1, sizeof(float) * 16 * S, cudaMemcpyDeviceToDevice, sb));
    rc = orbx_launch_pose_opt_slices(t->ctx, sb, S, D.estart, D.ecount, D.exw, D.eobs, D.eisg, &t->cam, D.T2, D.eoutlier,
                                     D.ninl + S, D.iters + 4 * S, K.d_scratch);
  }
  if (rc != ORBX_OK) return rc;
  TRK_EV(6, sb);
  ORBX_CUDA(cudaMemcpyAsync(d_Tcw_out, D.T2, sizeof(float) * 16 * S, cudaMemcpyDeviceToDevice, sb));
  if (t->chain) ORBX_CUDA(cudaMemcpyAsync(t->d_Tlast, D.T2, sizeof(float) * 16 * S, cudaMemcpyDeviceToDevice, sb));
  if (d_stats) {
    stats_kernel<<<S, 256, 0, sb>>>(D);
    ORBX_LAUNCH(t->ctx);
    ORBX_CUDA(cudaMemcpyAsync(d_stats, D.stats, sizeof(int) * ORBX_TRACK_STATS * S, cudaMemcpyDeviceToDevice, sb));
  }
  if (overlap) {
    ORBX_CUDA(cudaEventRecord(K.evB, sb));
    K.usedB = true;
  }
  if (t->kfPeriod > 0 && (t->stepCount + 1) % (unsigned long long)t->kfPeriod == 0) {
    if (t->kfTri) { rc = orbx_tri_batch_run(t->kfTri, t->stK); if (rc != ORBX_OK) return rc; }
    if (t->kfLba) { rc = orbx_lba_batch_run(t->kfLba, t->stK); if (rc != ORBX_OK) return rc; }
    t->kfRuns++;
  }
  t->stepCount++;
  t->profiled = t->profiling;
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

static void tracker_drop_graph(orbx_tracker* t) {
  if (t->graphExec) cudaGraphExecDestroy(t->graphExec);
  t->graphExec = nullptr;
  t->graphSeen = orbx_tracker::GraphKey{};
}

// One CUDA graph for the whole step.  A step whose arguments (device pointers, geometry, map binding) equal those of the
// previous call is captured once with stream capture of the very same code path and replayed with ONE cudaGraphLaunch from
// then on: the ~45 launches of a single frame's chain are mostly a few microseconds long, so the host's launch rate is what
// separates them.  Only the plain single-stream configuration is captured (no overlap mode, profiling, keyframe work or
// inertial upload events); anything else, and any change of the arguments, runs eagerly and drops the graph.
int orbx_tracker_set_graph(orbx_tracker* t, int enable) {
  if (!t) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  int rc = orbx_tracker_synchronize(t);
  if (rc != ORBX_OK) return rc;
  tracker_drop_graph(t);
  t->graphMode = enable != 0;
  return ORBX_OK;
}
long long orbx_tracker_graph_launches(const orbx_tracker* t) { return t ? (long long)t->graphLaunches : -1; }

int orbx_tracker_step_device(orbx_tracker* t, const uint8_t* d_imgs, int w, int h, int stride, const float* d_Tcw_true,
                             const float* d_Tcw_prior, float* d_Tcw_out, int32_t* d_stats) {
  if (!t || !d_imgs || !d_Tcw_true || !d_Tcw_prior || !d_Tcw_out) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  const bool plain = t->graphMode && t->stB == t->stA && !t->profiling && t->kfPeriod == 0 && !t->imu.mode;
  if (!plain) {
    if (t->graphExec) tracker_drop_graph(t);
    return tracker_step_body(t, d_imgs, w, h, stride, d_Tcw_true, d_Tcw_prior, d_Tcw_out, d_stats);
  }
  const orbx_tracker::GraphKey key{d_imgs, d_Tcw_true, d_Tcw_prior, d_Tcw_out, d_stats, w, h, stride, t->mapVersion, t->chain};
  if (t->graphExec && key == t->graphKey) {
    ORBX_CUDA(cudaGraphLaunch(t->graphExec, t->stA));
    t->ctx->launches.fetch_add(t->graphKernels, std::memory_order_relaxed);
    t->graphLaunches++;
    t->stepCount++;
    return ORBX_OK;
  }
  if (t->graphExec) tracker_drop_graph(t);
  // capture only a step that would not rebind anything: same arguments as the previous (eager) call, map already bound
  const bool bound = t->slot[0].mapVersion == t->mapVersion && w == t->argW && h == t->argH && stride == t->argStride && d_imgs == t->argImgs;
  if (!(key == t->graphSeen) || !bound) {
    t->graphSeen = key;
    return tracker_step_body(t, d_imgs, w, h, stride, d_Tcw_true, d_Tcw_prior, d_Tcw_out, d_stats);
  }
  cudaGraph_t graph = nullptr;
  const unsigned long long launches0 = t->ctx->launches.load(std::memory_order_relaxed), step0 = t->stepCount;
  if (cudaStreamBeginCapture(t->stA, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return tracker_step_body(t, d_imgs, w, h, stride, d_Tcw_true, d_Tcw_prior, d_Tcw_out, d_stats);
  }
  const int rc = tracker_step_body(t, d_imgs, w, h, stride, d_Tcw_true, d_Tcw_prior, d_Tcw_out, d_stats);
  const cudaError_t ce = cudaStreamEndCapture(t->stA, &graph);
  t->stepCount = step0;                                  // nothing has run yet
  const size_t nk = (size_t)(t->ctx->launches.load(std::memory_order_relaxed) - launches0);
  t->ctx->launches.store(launches0, std::memory_order_relaxed);
  if (rc != ORBX_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&t->graphExec, graph, 0) != cudaSuccess) {
    // not capturable after all (or the step failed): forget the graph path for these arguments and run eagerly
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    t->graphExec = nullptr;
    t->graphMode = false;
    return tracker_step_body(t, d_imgs, w, h, stride, d_Tcw_true, d_Tcw_prior, d_Tcw_out, d_stats);
  }
  cudaGraphDestroy(graph);
  t->graphKey = key;
  t->graphKernels = nk;
  ORBX_CUDA(cudaGraphLaunch(t->graphExec, t->stA));
  t->ctx->launches.fetch_add(nk, std::memory_order_relaxed);
  t->graphLaunches++;
  t->stepCount++;
  return ORBX_OK;
}

int orbx_tracker_step(orbx_tracker* t, const uint8_t* const* imgs, int w, int h, int stride, const float* Tcw_true,
                      const float* Tcw_prior, float* Tcw_out, int32_t* stats) {
  if (!t || !imgs || !Tcw_true || !Tcw_prior || !Tcw_out || w <= 0 || h <= 0 || stride < w) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  const int S = t->S;
  const size_t img = (size_t)w * h;
  if (!t->h_imgs) {
    ORBX_CUDA(cudaMallocHost(&t->h_imgs, (size_t)t->ips * S * img));
    ORBX_CUDA(cudaMallocHost(&t->h_pose, sizeof(float) * 16 * S * 3));
    ORBX_CUDA(cudaMallocHost(&t->h_stats, sizeof(int) * ORBX_TRACK_STATS * S));
    t->d_imgs = talloc<uint8_t>(t, (size_t)t->ips * S * img);
    t->d_hposeIn = talloc<float>(t, 32 * S);
    t->d_hposeOut = talloc<float>(t, 16 * S);
    t->d_hstats = talloc<int>(t, ORBX_TRACK_STATS * S);
    if (!t->d_imgs || !t->d_hposeIn || !t->d_hposeOut || !t->d_hstats) return ORBX_ECUDA;
  }
  cudaStream_t sa = t->stA, sb = t->stB;
  // page-locked caller memory is DMA'd directly (one copy when the images are also contiguous); anything else is
  // staged through the tracker's own pinned buffer first
  bool pinned = true, contiguous = stride == w;
  for (int b = 0; b < t->ips * S && pinned; ++b) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, imgs[b]) != cudaSuccess || at.type != cudaMemoryTypeHost) pinned = false;
    if (b > 0 && imgs[b] != imgs[b - 1] + img) contiguous = false;
  }
  cudaGetLastError();   // a failed attribute query on pageable memory leaves a sticky-free error behind
  if (pinned && contiguous) {
    ORBX_CUDA(cudaMemcpyAsync(t->d_imgs, imgs[0], (size_t)t->ips * S * img, cudaMemcpyHostToDevice, sa));
  } else if (pinned) {
    for (int b = 0; b < t->ips * S; ++b)
      ORBX_CUDA(cudaMemcpy2DAsync(t->d_imgs + b * img, w, imgs[b], stride, w, h, cudaMemcpyHostToDevice, sa));
  } else {
    for (int b = 0; b < t->ips * S; ++b)
      for (int y = 0; y < h; ++y) memcpy(t->h_imgs + b * img + (size_t)y * w, imgs[b] + (size_t)y * stride, w);
    ORBX_CUDA(cudaMemcpyAsync(t->d_imgs, t->h_imgs, (size_t)t->ips * S * img, cudaMemcpyHostToDevice, sa));
  }
  memcpy(t->h_pose, Tcw_true, sizeof(float) * 16 * S);
  memcpy(t->h_pose + 16 * S, Tcw_prior, sizeof(float) * 16 * S);
  ORBX_CUDA(cudaMemcpyAsync(t->d_hposeIn, t->h_pose, sizeof(float) * 32 * S, cudaMemcpyHostToDevice, sa));
  int rc = orbx_tracker_step_device(t, t->d_imgs, w, h, w, t->d_hposeIn, t->d_hposeIn + 16 * S, t->d_hposeOut, t->d_hstats);
  if (rc != ORBX_OK) return rc;
  ORBX_CUDA(cudaMemcpyAsync(t->h_pose + 32 * S, t->d_hposeOut, sizeof(float) * 16 * S, cudaMemcpyDeviceToHost, sb));
  ORBX_CUDA(cudaMemcpyAsync(t->h_stats, t->d_hstats, sizeof(int) * ORBX_TRACK_STATS * S, cudaMemcpyDeviceToHost, sb));
  ORBX_CUDA(cudaStreamSynchronize(sb));
  if (sb != sa) ORBX_CUDA(cudaStreamSynchronize(sa));
  memcpy(Tcw_out, t->h_pose + 32 * S, sizeof(float) * 16 * S);
  if (stats) memcpy(stats, t->h_stats, sizeof(int) * ORBX_TRACK_STATS * S);
  return ORBX_OK;
}

// Asynchronous host-buffer pipeline: submit() only enqueues — pinned H2D of the 2*S images on a copy stream into a staging
// buffer, one device-to-device copy of the staging buffer into the pyramid's level 0 at the head of stage A (which
// frees the staging buffer for the NEXT submit's H2D while this step's kernels run), the whole step in overlap mode,
// and the D2H of poses + statistics behind stage B.  collect() waits for the oldest outstanding submit.  At most two
// submits may be outstanding (the tracker is double-buffered).
int orbx_tracker_submit(orbx_tracker* t, const uint8_t* const* imgs, int w, int h, int stride, const float* Tcw_true,
                        const float* Tcw_prior) {
  if (!t || !imgs || !Tcw_true || !Tcw_prior || w <= 0 || h <= 0 || stride < w) return ORBX_EINVAL;
  if (t->submitted - t->collected >= 2) {
    orbx_set_error("orbx_tracker_submit: two steps are already outstanding; collect one first");
    return ORBX_ECAP;
  }
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  const int S = t->S;
  const size_t img = (size_t)w * h;
  if (t->stB == t->stA) {
    int rc = orbx_tracker_set_overlap(t, 1);
    if (rc != ORBX_OK) return rc;
  }
  if (!t->stC) {
    ORBX_CUDA(cudaStreamCreateWithFlags(&t->stC, cudaStreamNonBlocking));
    ORBX_CUDA(cudaEventCreateWithFlags(&t->evH2D, cudaEventDisableTiming));
    ORBX_CUDA(cudaEventCreateWithFlags(&t->evStaged, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) {
      ORBX_CUDA(cudaEventCreateWithFlags(&t->evDone[k], cudaEventDisableTiming));
      ORBX_CUDA(cudaMallocHost(&t->h_aPoseIn[k], sizeof(float) * 32 * S));
      ORBX_CUDA(cudaMallocHost(&t->h_aPoseOut[k], sizeof(float) * 16 * S));
      ORBX_CUDA(cudaMallocHost(&t->h_aStats[k], sizeof(int) * ORBX_TRACK_STATS * S));
      t->d_aPoseIn[k] = talloc<float>(t, 32 * S);
      t->d_aPoseOut[k] = talloc<float>(t, 16 * S);
      t->d_aStats[k] = talloc<int>(t, ORBX_TRACK_STATS * S);
      if (!t->d_aPoseIn[k] || !t->d_aPoseOut[k] || !t->d_aStats[k]) return ORBX_ECUDA;
    }
  }
  if (!t->h_imgs) {
    ORBX_CUDA(cudaMallocHost(&t->h_imgs, (size_t)t->ips * S * img));
    ORBX_CUDA(cudaMallocHost(&t->h_pose, sizeof(float) * 16 * S * 3));
    ORBX_CUDA(cudaMallocHost(&t->h_stats, sizeof(int) * ORBX_TRACK_STATS * S));
    t->d_imgs = talloc<uint8_t>(t, (size_t)t->ips * S * img);
    t->d_hposeIn = talloc<float>(t, 32 * S);
    t->d_hposeOut = talloc<float>(t, 16 * S);
    t->d_hstats = talloc<int>(t, ORBX_TRACK_STATS * S);
    if (!t->d_imgs || !t->d_hposeIn || !t->d_hposeOut || !t->d_hstats) return ORBX_ECUDA;
  }
  size_t l0bytes = 0;
  uint8_t* level0 = orbx_ext_level0_storage(t->ext, &l0bytes);
  if (!level0 || l0bytes < (size_t)t->ips * S * img) {
    orbx_set_error("orbx_tracker_submit: extractor level-0 storage too small for %d images of %dx%d", t->ips * S, w, h);
    return ORBX_ECAP;
  }
  const int k = (int)(t->stepCount & 1);            // the slot orbx_tracker_step_device is about to use
  cudaStream_t sa = t->stA, sb = t->stB, sc = t->stC;
  // ---- copy stream: inputs of this step ----
  if (t->stagedUsed) ORBX_CUDA(cudaStreamWaitEvent(sc, t->evStaged, 0));   // staging buffer consumed by the previous step
  bool pinned = true, contiguous = stride == w;
  for (int b = 0; b < t->ips * S && pinned; ++b) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, imgs[b]) != cudaSuccess || at.type != cudaMemoryTypeHost) pinned = false;
    if (b > 0 && imgs[b] != imgs[b - 1] + img) contiguous = false;
  }
  cudaGetLastError();
  if (pinned && contiguous) {
    ORBX_CUDA(cudaMemcpyAsync(t->d_imgs, imgs[0], (size_t)t->ips * S * img, cudaMemcpyHostToDevice, sc));
  } else if (pinned) {
    for (int b = 0; b < t->ips * S; ++b)
      ORBX_CUDA(cudaMemcpy2DAsync(t->d_imgs + b * img, w, imgs[b], stride, w, h, cudaMemcpyHostToDevice, sc));
  } else {
    if (t->submitted) ORBX_CUDA(cudaEventSynchronize(t->evH2D));          // the pinned bounce buffer is free again
    for (int b = 0; b < t->ips * S; ++b)
      for (int y = 0; y < h; ++y) memcpy(t->h_imgs + b * img + (size_t)y * w, imgs[b] + (size_t)y * stride, w);
    ORBX_CUDA(cudaMemcpyAsync(t->d_imgs, t->h_imgs, (size_t)t->ips * S * img, cudaMemcpyHostToDevice, sc));
  }
  memcpy(t->h_aPoseIn[k], Tcw_true, sizeof(float) * 16 * S);
  memcpy(t->h_aPoseIn[k] + 16 * S, Tcw_prior, sizeof(float) * 16 * S);
  ORBX_CUDA(cudaMemcpyAsync(t->d_aPoseIn[k], t->h_aPoseIn[k], sizeof(float) * 32 * S, cudaMemcpyHostToDevice, sc));
  ORBX_CUDA(cudaEventRecord(t->evH2D, sc));
  // ---- stage A stream: staging -> level 0, then the step ----
  ORBX_CUDA(cudaStreamWaitEvent(sa, t->evH2D, 0));
  ORBX_CUDA(cudaMemcpyAsync(level0, t->d_imgs, (size_t)t->ips * S * img, cudaMemcpyDeviceToDevice, sa));
  ORBX_CUDA(cudaEventRecord(t->evStaged, sa));
  t->stagedUsed = true;
  int rc = orbx_tracker_step_device(t, level0, w, h, w, t->d_aPoseIn[k], t->d_aPoseIn[k] + 16 * S, t->d_aPoseOut[k], t->d_aStats[k]);
  if (rc != ORBX_OK) return rc;
  sb = t->stB;
  ORBX_CUDA(cudaMemcpyAsync(t->h_aPoseOut[k], t->d_aPoseOut[k], sizeof(float) * 16 * S, cudaMemcpyDeviceToHost, sb));
  ORBX_CUDA(cudaMemcpyAsync(t->h_aStats[k], t->d_aStats[k], sizeof(int) * ORBX_TRACK_STATS * S, cudaMemcpyDeviceToHost, sb));
  ORBX_CUDA(cudaEventRecord(t->evDone[k], sb));
  t->ringSlot[t->submitted & 1] = k;
  t->submitted++;
  return ORBX_OK;
}

int orbx_tracker_collect(orbx_tracker* t, float* Tcw_out, int32_t* stats) {
  if (!t || !Tcw_out) return ORBX_EINVAL;
  if (t->collected == t->submitted) {
    orbx_set_error("orbx_tracker_collect: nothing outstanding");
    return ORBX_EINVAL;
  }
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  const int k = t->ringSlot[t->collected & 1];
  ORBX_CUDA(cudaEventSynchronize(t->evDone[k]));
  memcpy(Tcw_out, t->h_aPoseOut[k], sizeof(float) * 16 * t->S);
  if (stats) memcpy(stats, t->h_aStats[k], sizeof(int) * ORBX_TRACK_STATS * t->S);
  t->collected++;
  return ORBX_OK;
}

int orbx_tracker_set_profiling(orbx_tracker* t, int enable) {
  if (!t) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  if (enable && !t->ev[0])
    for (int i = 0; i <= ORBX_TRACK_STAGES; ++i) ORBX_CUDA(cudaEventCreate(&t->ev[i]));
  t->profiling = enable != 0;
  t->profiled = false;
  return ORBX_OK;
}

int orbx_tracker_stage_ms(orbx_tracker* t, float* ms) {
  if (!t || !ms || !t->profiled) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(t->ctx->device));
  ORBX_CUDA(cudaEventSynchronize(t->ev[ORBX_TRACK_STAGES]));
  for (int i = 0; i < ORBX_TRACK_STAGES; ++i) ORBX_CUDA(cudaEventElapsedTime(&ms[i], t->ev[i], t->ev[i + 1]));
  return ORBX_OK;
}

}  // extern "C"
