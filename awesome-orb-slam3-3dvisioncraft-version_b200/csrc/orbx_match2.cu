// orbx_match2.cu — the matchers either side of the hot path: ORBmatcher::SearchByBoW(KeyFrame*, Frame&)
// src/ORBmatcher.cc:323-591 and the search half of ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>, th, bRight)
// src/ORBmatcher.cc:1630-1883 (with MapPoint::PredictScale src/MapPoint.cc:578-593, KeyFrame::GetFeaturesInArea /
// IsInImage src/KeyFrame.cc:810-859, Pinhole::project src/CameraModels/Pinhole.cpp:31-50).
//
//   K15 bow_match_kernel : one warp per vocabulary node shared by the two FeatureVectors.  Inside a node the reference is
//       sequential (a frame keypoint taken by an earlier keyframe feature is skipped by later ones), so the warp walks the
//       keyframe's features in order and spends its 32 lanes on the frame's features of that node (__popc Hamming,
//       best / second best by two warp min-reductions).  Different nodes never share a keypoint, so nodes run in parallel.
//   K16 bow_rot_filter_kernel : rotation histogram + ComputeThreeMaxima + removal, one CTA.
//   K17 fuse_kernel : one warp per MapPoint: project, distance / viewing-angle / scale gates, grid walk in the
//       reference's order, chi2 gates, first minimum Hamming distance.
#include "orbx_match.cuh"

#define TH_LOW 50
#define HISTO_LENGTH 30

__device__ __forceinline__ int m2_hamming256(const uint4* __restrict__ a, const uint4* __restrict__ b) {
  const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ __forceinline__ int m2_rot_bin(float a, float b) {
  const float factor = 1.0f / HISTO_LENGTH;
  float rot = __fsub_rn(a, b);
  if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
  int bin = (int)roundf(__fmul_rn(rot, factor));
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

// ComputeThreeMaxima (src/ORBmatcher.cc:2654-2695) on bin counts
__device__ void m2_three_maxima(const int* h, int& i1, int& i2, int& i3) {
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

struct BowArgs {
  int nK, nF;                      // keypoints of the keyframe / frame
  const orbx_keypoint *kpK, *kpF;
  const uint8_t *descK, *descF;
  const uint8_t* hasMp;            // [nK]
  int nnK, nnF;
  const int *kNode, *kOff, *kIdx, *fNode, *fOff, *fIdx;
  float nnratio;
  int checkOri;
  volatile int* matchF;            // [nF] keyframe feature assigned to frame keypoint i, or -1
  int* nmatches;
};

__global__ void __launch_bounds__(128) bow_match_kernel(BowArgs A) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (a >= A.nnK) return;
  const int node = A.kNode[a];
  int lo = 0, hi = A.nnF;                                   // lower_bound of the node id in the frame's FeatureVector
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.fNode[mid] < node) lo = mid + 1; else hi = mid; }
  if (lo >= A.nnF || A.fNode[lo] != node) return;
  const int fb = A.fOff[lo], fe = A.fOff[lo + 1];
  const unsigned FULL = 0xffffffffu;
  for (int iK = A.kOff[a]; iK < A.kOff[a + 1]; ++iK) {
    const int realIdxKF = A.kIdx[iK];
    if (!A.hasMp[realIdxKF]) continue;
    const uint4* dK = reinterpret_cast<const uint4*>(A.descK + 32 * (size_t)realIdxKF);
    unsigned k1 = 0xffffffffu, k2 = 0xffffffffu;            // two smallest keys, key = dist << 20 | position
    for (int c0 = fb; c0 < fe; c0 += 32) {
      const int i = c0 + lane;
      unsigned key = 0xffffffffu;
      if (i < fe) {
        const int idxF = A.fIdx[i];
        if (A.matchF[idxF] < 0)                             // vpMapPointMatches[realIdxF] == NULL
          key = ((unsigned)m2_hamming256(dK, reinterpret_cast<const uint4*>(A.descF + 32 * (size_t)idxF)) << 20) | (unsigned)(i - fb);
      }
      const unsigned m1 = __reduce_min_sync(FULL, key);
      const unsigned m2 = __reduce_min_sync(FULL, key == m1 ? 0xffffffffu : key);
      if (m1 < k1) { k2 = min(k1, m2); k1 = m1; }
      else { k2 = min(k2, m1); }
    }
    if (k1 != 0xffffffffu) {
      const int best1 = k1 >> 20, best2 = k2 == 0xffffffffu ? 256 : (int)(k2 >> 20);
      if (best1 <= TH_LOW && (float)best1 < __fmul_rn(A.nnratio, (float)best2)) {
        if (lane == 0) A.matchF[A.fIdx[fb + (k1 & 0xfffff)]] = realIdxKF;
      }
    }
    __syncwarp();                                           // the assignment is visible to the next feature's scan
  }
}

__global__ void __launch_bounds__(256) bow_rot_filter_kernel(BowArgs A) {
  __shared__ int s_hist[HISTO_LENGTH];
  __shared__ int s_keep[3];
  __shared__ int s_n;
  if (threadIdx.x < HISTO_LENGTH) s_hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < A.nF; i += 256) {
    const int m = A.matchF[i];
    if (m < 0) continue;
    ++local;
    if (A.checkOri) atomicAdd(&s_hist[m2_rot_bin(A.kpK[m].angle, A.kpF[i].angle)], 1);
  }
  if (local) atomicAdd(&s_n, local);
  __syncthreads();
  if (A.checkOri) {
    if (threadIdx.x == 0) { int i1, i2, i3; m2_three_maxima(s_hist, i1, i2, i3); s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3; }
    __syncthreads();
    int removed = 0;
    for (int i = threadIdx.x; i < A.nF; i += 256) {
      const int m = A.matchF[i];
      if (m < 0) continue;
      const int bin = m2_rot_bin(A.kpK[m].angle, A.kpF[i].angle);
      if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { A.matchF[i] = -1; ++removed; }
    }
    if (removed) atomicSub(&s_n, removed);
    __syncthreads();
  }
  if (threadIdx.x == 0) *A.nmatches = s_n;
}

// ------------------------------------------------------------------------------------
// Fuse
// ------------------------------------------------------------------------------------
struct FuseArgs {
  int nmp;
  const uint8_t* flags;
  const float *xw, *maxDist, *minDist, *normal;
  const uint8_t* mpDesc;
  float R[9], t[3], Ow[3];
  float fx, fy, cx, cy, bf;
  float th, logScale;
  int nlevels;
  float scaleFactors[ORBX_MAX_LEVELS], invSigma2[ORBX_MAX_LEVELS];
  int* bestIdx;
  int* nfused;
};

// GetFeaturesInArea cell window (src/KeyFrame.cc:818-832; same arithmetic as Frame's)
__device__ __forceinline__ bool m2_cell_range(const FrameDev& F, float x, float y, float r, int& x0, int& x1, int& y0, int& y1) {
  x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, F.minX), r), F.wInv)));
  if (x0 >= ORBX_GRID_COLS) return false;
  x1 = min(ORBX_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, F.minX), r), F.wInv)));
  if (x1 < 0) return false;
  y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, F.minY), r), F.hInv)));
  if (y0 >= ORBX_GRID_ROWS) return false;
  y1 = min(ORBX_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, F.minY), r), F.hInv)));
  if (y1 < 0) return false;
  return true;
}

__global__ void __launch_bounds__(128) fuse_kernel(const FrameDev* frames, FuseArgs A) {
  const FrameDev F = frames[0];
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= A.nmp) return;
  if (lane == 0) A.bestIdx[q] = -1;
  if (!(A.flags[q] & 1)) return;
  const float X = A.xw[3 * q], Y = A.xw[3 * q + 1], Z = A.xw[3 * q + 2];
  // p3Dc = Rcw*p3Dw + tcw in fp32, fixed left-to-right order
  const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A.R[0], X), __fmul_rn(A.R[1], Y)), __fmul_rn(A.R[2], Z)), A.t[0]);
  const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A.R[3], X), __fmul_rn(A.R[4], Y)), __fmul_rn(A.R[5], Z)), A.t[1]);
  const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(A.R[6], X), __fmul_rn(A.R[7], Y)), __fmul_rn(A.R[8], Z)), A.t[2]);
  if (zc < 0.0f) return;
  const float invz = __fdiv_rn(1.0f, zc);
  const float u = __fadd_rn(__fdiv_rn(__fmul_rn(A.fx, xc), zc), A.cx), v = __fadd_rn(__fdiv_rn(__fmul_rn(A.fy, yc), zc), A.cy);
  if (!(u >= F.minX && u < F.maxX && v >= F.minY && v < F.maxY)) return;
  const float ur = __fsub_rn(u, __fmul_rn(A.bf, invz));
  const float maxDistance = __fmul_rn(1.2f, A.maxDist[q]), minDistance = __fmul_rn(0.8f, A.minDist[q]);
  const float P0 = __fsub_rn(X, A.Ow[0]), P1 = __fsub_rn(Y, A.Ow[1]), P2 = __fsub_rn(Z, A.Ow[2]);
  const float dist3D = (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn((double)P0, (double)P0), __dmul_rn((double)P1, (double)P1)),
                                             __dmul_rn((double)P2, (double)P2)));
  if (dist3D < minDistance || dist3D > maxDistance) return;
  const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)P0, (double)A.normal[3 * q]), __dmul_rn((double)P1, (double)A.normal[3 * q + 1])),
                               __dmul_rn((double)P2, (double)A.normal[3 * q + 2]));
  if (dot < __dmul_rn(0.5, (double)dist3D)) return;
  const float ratio = __fdiv_rn(A.maxDist[q], dist3D);
  int level = (int)ceil(log((double)ratio) / (double)A.logScale);   // MapPoint::PredictScale
  if (level < 0) level = 0; else if (level >= A.nlevels) level = A.nlevels - 1;
  const float radius = __fmul_rn(A.th, A.scaleFactors[level]);
  int x0, x1, y0, y1;
  if (!m2_cell_range(F, u, v, radius, x0, x1, y0, y1)) return;
  const uint4* dMP = reinterpret_cast<const uint4*>(A.mpDesc + 32 * (size_t)q);
  unsigned key = 0xffffffffu;   // dist << 20 | visiting position: the first minimum wins, like `dist < bestDist`
  int keyIdx = -1;
  int pos = 0;
  for (int ix = x0; ix <= x1; ++ix) {
    const int beg = F.cellStart[ix * ORBX_GRID_ROWS + y0], end = F.cellStart[ix * ORBX_GRID_ROWS + y1 + 1];
    for (int base = beg; base < end; base += 32) {
      const int i = base + lane;
      bool inArea = false, ok = false;
      int idx = -1;
      orbx_keypoint kp;
      if (i < end) {
        idx = F.cellIdx[i];
        kp = F.kps[idx];
        inArea = fabsf(__fsub_rn(kp.x, u)) < radius && fabsf(__fsub_rn(kp.y, v)) < radius;
      }
      const unsigned m = __ballot_sync(0xffffffffu, inArea);
      if (inArea) {
        const int kpLevel = kp.octave;
        ok = !(kpLevel < level - 1 || kpLevel > level);
        if (ok) {
          const float ex = __fsub_rn(u, kp.x), ey = __fsub_rn(v, kp.y);
          const float urk = F.uright ? F.uright[idx] : -1.0f;
          if (urk >= 0) {
            const float er = __fsub_rn(ur, urk);
            const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
            ok = !((double)__fmul_rn(e2, A.invSigma2[kpLevel]) > 7.8);
          } else {
            const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
            ok = !((double)__fmul_rn(e2, A.invSigma2[kpLevel]) > 5.99);
          }
        }
        if (ok) {
          const unsigned p = (unsigned)(pos + __popc(m & ((1u << lane) - 1)));
          const unsigned k = ((unsigned)m2_hamming256(dMP, reinterpret_cast<const uint4*>(F.desc + 32 * (size_t)idx)) << 20) | p;
          if (k < key) { key = k; keyIdx = idx; }
        }
      }
      pos += __popc(m);
    }
  }
  const unsigned best = __reduce_min_sync(0xffffffffu, key);
  if (best == 0xffffffffu || (int)(best >> 20) > TH_LOW) return;
  if (key == best) {            // exactly one lane owns the winning (distance, position) pair
    A.bestIdx[q] = keyIdx;
    atomicAdd(A.nfused, 1);
  }
}

extern "C" {

int orbx_search_by_bow(orbx_ctx* ctx, const orbx_frame_desc* kf, const orbx_frame_desc* frame, const uint8_t* kf_has_mp,
                       int nn_kf, const int32_t* fv_kf_node, const int32_t* fv_kf_off, const int32_t* fv_kf_idx, int nn_f,
                       const int32_t* fv_f_node, const int32_t* fv_f_off, const int32_t* fv_f_idx, float nnratio,
                       int check_orientation, int32_t* match_f, int32_t* nmatches) {
  if (!ctx || !kf || !frame || !kf_has_mp || nn_kf < 0 || nn_f < 0 || !match_f || !nmatches) return ORBX_EINVAL;
  if ((nn_kf && (!fv_kf_node || !fv_kf_off || !fv_kf_idx)) || (nn_f && (!fv_f_node || !fv_f_off || !fv_f_idx))) return ORBX_EINVAL;
  if (kf->n < 0 || frame->n < 0 || (kf->n && (!kf->kps || !kf->desc)) || (frame->n && (!frame->kps || !frame->desc))) return ORBX_EINVAL;
  *nmatches = 0;
  for (int i = 0; i < frame->n; ++i) match_f[i] = -1;
  const int totK = nn_kf ? fv_kf_off[nn_kf] : 0, totF = nn_f ? fv_f_off[nn_f] : 0;
  if (frame->n == 0 || kf->n == 0 || totK == 0 || totF == 0) return ORBX_OK;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  BowArgs A;
  A.nK = kf->n; A.nF = frame->n;
  A.kpK = S.upload(kf->kps, kf->n); A.kpF = S.upload(frame->kps, frame->n);
  A.descK = S.upload(kf->desc, (size_t)kf->n * 32); A.descF = S.upload(frame->desc, (size_t)frame->n * 32);
  A.hasMp = S.upload(kf_has_mp, kf->n);
  A.nnK = nn_kf; A.nnF = nn_f;
  A.kNode = S.upload(fv_kf_node, nn_kf); A.kOff = S.upload(fv_kf_off, nn_kf + 1); A.kIdx = S.upload(fv_kf_idx, totK);
  A.fNode = S.upload(fv_f_node, nn_f); A.fOff = S.upload(fv_f_off, nn_f + 1); A.fIdx = S.upload(fv_f_idx, totF);
  A.nnratio = nnratio;
  A.checkOri = check_orientation;
  int* dMatch = S.alloc<int>(frame->n);
  A.matchF = dMatch;
  A.nmatches = S.alloc<int>(1);
  if (S.failed) return ORBX_ECUDA;
  ORBX_CUDA(cudaMemsetAsync(dMatch, 0xff, sizeof(int) * frame->n, st));
  bow_match_kernel<<<div_up(nn_kf * 32, 128), 128, 0, st>>>(A);
  ORBX_LAUNCH(ctx);
  bow_rot_filter_kernel<<<1, 256, 0, st>>>(A);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  S.download(match_f, (const int32_t*)dMatch, (size_t)frame->n);
  S.download(nmatches, (const int32_t*)A.nmatches, (size_t)1);
  return S.finish();
}

int orbx_fuse(orbx_ctx* ctx, const orbx_frame_desc* kf, const orbx_camera* cam, const float* Rcw, const float* tcw,
              const float* Ow, int nmp, const uint8_t* flags, const float* xw, const float* mp_max_dist,
              const float* mp_min_dist, const float* mp_normal, const uint8_t* mp_desc, float th,
              const float* scale_factors, const float* inv_level_sigma2, int nlevels, float log_scale_factor,
              int32_t* best_idx, int32_t* nfused) {
  if (!ctx || !kf || !cam || !Rcw || !tcw || !Ow || nmp < 0 || !best_idx || !nfused || !scale_factors || !inv_level_sigma2 ||
      nlevels < 1 || nlevels > ORBX_MAX_LEVELS)
    return ORBX_EINVAL;
  if (nmp && (!flags || !xw || !mp_max_dist || !mp_min_dist || !mp_normal || !mp_desc)) return ORBX_EINVAL;
  *nfused = 0;
  for (int i = 0; i < nmp; ++i) best_idx[i] = -1;
  if (nmp == 0 || kf->n == 0) return ORBX_OK;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  FrameDev F;
  int rc = orbx_upload_frame(S, kf, &F);
  if (rc != ORBX_OK) return rc;
  FrameDev* dF = S.upload(&F, 1);
  FuseArgs A;
  A.nmp = nmp;
  A.flags = S.upload(flags, nmp);
  A.xw = S.upload(xw, (size_t)3 * nmp);
  A.maxDist = S.upload(mp_max_dist, nmp);
  A.minDist = S.upload(mp_min_dist, nmp);
  A.normal = S.upload(mp_normal, (size_t)3 * nmp);
  A.mpDesc = S.upload(mp_desc, (size_t)32 * nmp);
  for (int i = 0; i < 9; ++i) A.R[i] = Rcw[i];
  for (int i = 0; i < 3; ++i) { A.t[i] = tcw[i]; A.Ow[i] = Ow[i]; }
  A.fx = cam->fx; A.fy = cam->fy; A.cx = cam->cx; A.cy = cam->cy; A.bf = cam->bf;
  A.th = th;
  A.logScale = log_scale_factor;
  A.nlevels = nlevels;
  for (int l = 0; l < ORBX_MAX_LEVELS; ++l) {
    A.scaleFactors[l] = l < nlevels ? scale_factors[l] : 0.f;
    A.invSigma2[l] = l < nlevels ? inv_level_sigma2[l] : 0.f;
  }
  A.bestIdx = S.alloc<int>(nmp);
  A.nfused = S.alloc<int>(1);
  if (S.failed) return ORBX_ECUDA;
  ORBX_CUDA(cudaMemsetAsync(A.nfused, 0, sizeof(int), st));
  if (!F.gridBuilt) {
    rc = orbx_launch_grid_build(ctx, st, dF, 1);
    if (rc != ORBX_OK) return rc;
  }
  fuse_kernel<<<div_up(nmp * 32, 128), 128, 0, st>>>(dF, A);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  S.download(best_idx, (const int32_t*)A.bestIdx, (size_t)nmp);
  S.download(nfused, (const int32_t*)A.nfused, (size_t)1);
  return S.finish();
}

}  // extern "C"
