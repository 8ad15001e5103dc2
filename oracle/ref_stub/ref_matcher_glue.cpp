// ref_matcher_glue.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's UNMODIFIED ORB_SLAM3::ORBmatcher
// (src/ORBmatcher.cc compiled in place behind matcher_prelude.h).  Every function takes exactly the flat arguments of
// the oracle's ork_* counterpart (= the C ABI of include/orbx.h), builds the stand-in Frame / KeyFrame / MapPoint objects
// from them, calls the reference member function and reads the result back out of the objects the reference mutated.
#include "ORBmatcher.h"   // the reference's own header; -include matcher_prelude.h has pre-empted its heavy includes
#include "../../include/orbx.h"
#include <cstring>
#include <memory>

extern "C" {
void* ork_grid_create(const orbx_frame_desc* F);
void ork_grid_destroy(void* g);
int ork_grid_query(const void* g, float x, float y, float r, int minL, int maxL, int32_t* out, int cap);
}

namespace ORB_SLAM3 {
struct RefGrid {
  std::vector<orbx_keypoint> kps;
  orbx_frame_desc fd;
  void* g;
};
RefGrid* ref_grid_build(const std::vector<cv::KeyPoint>& keysUn, float minX, float minY, float maxX, float maxY) {
  RefGrid* G = new RefGrid();
  G->kps.resize(keysUn.size());
  for (size_t i = 0; i < keysUn.size(); ++i)
    G->kps[i] = {keysUn[i].pt.x, keysUn[i].pt.y, keysUn[i].size, keysUn[i].angle, keysUn[i].response, keysUn[i].octave};
  std::memset(&G->fd, 0, sizeof(G->fd));
  G->fd.n = (int)keysUn.size();
  G->fd.kps = G->kps.data();
  G->fd.min_x = minX; G->fd.min_y = minY; G->fd.max_x = maxX; G->fd.max_y = maxY;
  G->g = ork_grid_create(&G->fd);
  return G;
}
void ref_grid_free(RefGrid* G) { ork_grid_destroy(G->g); delete G; }
std::vector<size_t> ref_grid_query(const RefGrid* G, const std::vector<cv::KeyPoint>&, float x, float y, float r, int minLevel,
                                   int maxLevel) {
  std::vector<int32_t> tmp(G->kps.size() + 1);
  const int n = ork_grid_query(G->g, x, y, r, minLevel, maxLevel, tmp.data(), (int)tmp.size());
  return std::vector<size_t>(tmp.begin(), tmp.begin() + n);
}

struct MatcherAccess : ORBmatcher {   // ComputeThreeMaxima is protected
  using ORBmatcher::ORBmatcher;
  void three(std::vector<int>* h, int L, int& a, int& b, int& c) { ComputeThreeMaxima(h, L, a, b, c); }
};
}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

static cv::Mat desc_row(const uint8_t* d) {
  cv::Mat m(1, 32, CV_8U);
  std::memcpy(m.ptr(), d, 32);
  return m;
}
static cv::Mat mat3x1(const float* v) { return (cv::Mat_<float>(3, 1) << v[0], v[1], v[2]); }
static cv::Mat mat_rows(const float* v, int r, int c) {
  cv::Mat m(r, c, CV_32F);
  for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = v[i * c + j];
  return m;
}
static std::vector<cv::KeyPoint> keys_of(const orbx_frame_desc* F) {
  std::vector<cv::KeyPoint> k(F->n);
  for (int i = 0; i < F->n; ++i) k[i] = cv::KeyPoint(F->kps[i].x, F->kps[i].y, F->kps[i].size, F->kps[i].angle, F->kps[i].response, F->kps[i].octave);
  return k;
}
static cv::Mat descs_of(const orbx_frame_desc* F) {
  cv::Mat m(std::max(F->n, 1), 32, CV_8U);
  if (F->n) std::memcpy(m.ptr(), F->desc, (size_t)32 * F->n);
  return m;
}
static void fill_levels(std::vector<float>& sf, std::vector<float>& s2, std::vector<float>& is2, const float* scaleFactors, int nlevels) {
  sf.assign(scaleFactors, scaleFactors + nlevels);
  s2.resize(nlevels); is2.resize(nlevels);
  for (int l = 0; l < nlevels; ++l) { s2[l] = sf[l] * sf[l]; is2[l] = 1.0f / s2[l]; }
}
static void fill_frame(Frame& F, const orbx_frame_desc* D, const float* scaleFactors, int nlevels) {
  F.N = D->n;
  F.mvKeysUn = keys_of(D);
  F.mvKeys = F.mvKeysUn;
  F.mDescriptors = descs_of(D);
  F.mvuRight.assign(D->n, -1.f);
  if (D->uright) F.mvuRight.assign(D->uright, D->uright + D->n);
  F.mvpMapPoints.assign(D->n, nullptr);
  F.mvbOutlier.assign(D->n, false);
  F.mnMinX = D->min_x; F.mnMinY = D->min_y; F.mnMaxX = D->max_x; F.mnMaxY = D->max_y;
  F.mnScaleLevels = nlevels;
  fill_levels(F.mvScaleFactors, F.mvLevelSigma2, F.mvInvLevelSigma2, scaleFactors, nlevels);
}
static void fill_keyframe(KeyFrame& K, const orbx_frame_desc* D, const float* scaleFactors, int nlevels) {
  K.N = D->n;
  K.mvKeysUn = keys_of(D);
  K.mvKeys = K.mvKeysUn;
  K.mDescriptors = descs_of(D);
  K.mvuRight.assign(D->n, -1.f);
  if (D->uright) K.mvuRight.assign(D->uright, D->uright + D->n);
  K.mvpMapPoints.assign(D->n, nullptr);
  K.mnMinX = (int)D->min_x; K.mnMinY = (int)D->min_y; K.mnMaxX = (int)D->max_x; K.mnMaxY = (int)D->max_y;
  K.mnScaleLevels = nlevels;
  if (scaleFactors) fill_levels(K.mvScaleFactors, K.mvLevelSigma2, K.mvInvLevelSigma2, scaleFactors, nlevels);
}
static void fill_featvec(DBoW2::FeatureVector& fv, int nn, const int32_t* node, const int32_t* off, const int32_t* idx) {
  for (int a = 0; a < nn; ++a) {
    std::vector<unsigned int> v(idx + off[a], idx + off[a + 1]);
    fv.insert(std::make_pair((DBoW2::NodeId)node[a], v));
  }
}

extern "C" {

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }

// ORBmatcher::ComputeThreeMaxima on a histogram given by its bin counts (src/ORBmatcher.cc:2654-2698)
void ref_three_maxima(const int32_t* counts, int L, int32_t* out3) {
  std::vector<std::vector<int>> h(L);
  for (int i = 0; i < L; ++i) h[i].assign(counts[i], 0);
  MatcherAccess m(0.6f, true);
  int a, b, c;
  m.three(h.data(), L, a, b, c);
  out3[0] = a; out3[1] = b; out3[2] = c;
}

// src/ORBmatcher.cc:59-255.  cur_mp_out[n]: final F.mvpMapPoints as MapPoint index (-1 none, -2 a pre-existing blocker)
int ref_search_by_projection_map(const orbx_frame_desc* Fd, const uint8_t* kp_blocked, int nq, const float* projX, const float* projY,
                                 const float* projXR, const int32_t* level, const float* viewCos, const uint8_t* mpDesc,
                                 const uint8_t* flags, float th, float nnratio, const float* scaleFactors, int nlevels,
                                 int32_t* cur_mp_out, int32_t* nmatches) {
  Frame F;
  fill_frame(F, Fd, scaleFactors, nlevels);
  MapPoint blocker;
  blocker.nObs = 1;
  for (int i = 0; i < Fd->n; ++i) if (kp_blocked[i]) F.mvpMapPoints[i] = &blocker;
  std::vector<MapPoint> mps(nq);
  std::vector<MapPoint*> vp(nq);
  for (int q = 0; q < nq; ++q) {
    MapPoint& m = mps[q];
    m.mbTrackInView = (flags[q] & 1) != 0;
    m.nObs = (flags[q] & 2) ? 1 : 0;
    m.mTrackProjX = projX[q]; m.mTrackProjY = projY[q]; m.mTrackProjXR = projXR[q];
    m.mnTrackScaleLevel = level[q]; m.mTrackViewCos = viewCos[q];
    m.mDescriptor = desc_row(mpDesc + (size_t)32 * q);
    vp[q] = &m;
  }
  ORBmatcher matcher(nnratio, true);
  *nmatches = matcher.SearchByProjection(F, vp, th, false, 50.0f);
  for (int i = 0; i < Fd->n; ++i) {
    MapPoint* p = F.mvpMapPoints[i];
    cur_mp_out[i] = !p ? -1 : (p == &blocker ? -2 : (int)(p - mps.data()));
  }
  return 0;
}

// src/ORBmatcher.cc:2244-2509.  cur_match[n]: final CurrentFrame.mvpMapPoints as last-frame index (-1 none, -2 blocker)
int ref_search_by_projection_frame(const orbx_frame_desc* Cd, const uint8_t* cur_blocked, const orbx_camera* cam, const float* Tc,
                                   const float* Tl, int nq, const uint8_t* flags, const float* xw, const int32_t* octave,
                                   const float* angle, const uint8_t* mpDesc, float th, int bMono, int checkOri,
                                   const float* scaleFactors, int nlevels, int32_t* cur_match, int32_t* nmatches) {
  Frame C, Lf;
  fill_frame(C, Cd, scaleFactors, nlevels);
  GeometricCamera camera(cam->fx, cam->fy, cam->cx, cam->cy);
  C.mpCamera = &camera;
  C.fx = cam->fx; C.fy = cam->fy; C.cx = cam->cx; C.cy = cam->cy; C.mbf = cam->bf; C.mb = cam->b;
  C.mTcw = mat_rows(Tc, 4, 4);
  MapPoint blocker;
  blocker.nObs = 1;
  for (int i = 0; i < Cd->n; ++i) if (cur_blocked[i]) C.mvpMapPoints[i] = &blocker;
  Lf.N = nq;
  Lf.mTcw = mat_rows(Tl, 4, 4);
  Lf.mvKeysUn.resize(nq);
  std::vector<MapPoint> mps(nq);
  Lf.mvpMapPoints.assign(nq, nullptr);
  Lf.mvbOutlier.assign(nq, false);
  for (int q = 0; q < nq; ++q) {
    Lf.mvKeysUn[q] = cv::KeyPoint(0, 0, 31, angle[q], 0, octave[q]);
    if (flags[q] & 1) {
      mps[q].mWorldPos = mat3x1(xw + 3 * q);
      mps[q].mDescriptor = desc_row(mpDesc + (size_t)32 * q);
      mps[q].nObs = (flags[q] & 2) ? 1 : 0;
      Lf.mvpMapPoints[q] = &mps[q];
    }
  }
  Lf.mvKeys = Lf.mvKeysUn;
  Lf.mpCamera = &camera;
  ORBmatcher matcher(0.9f, checkOri != 0);
  *nmatches = matcher.SearchByProjection(C, Lf, th, bMono != 0);
  for (int i = 0; i < Cd->n; ++i) {
    MapPoint* p = C.mvpMapPoints[i];
    cur_match[i] = !p ? -1 : (p == &blocker ? -2 : (int)(p - mps.data()));
  }
  return 0;
}

// src/ORBmatcher.cc:1138-1428
int ref_search_for_triangulation(const orbx_frame_desc* K1d, const orbx_frame_desc* K2d, const uint8_t* has1, const uint8_t* has2,
                                 int nn1, const int32_t* n1id, const int32_t* n1off, const int32_t* n1idx, int nn2,
                                 const int32_t* n2id, const int32_t* n2off, const int32_t* n2idx, const orbx_camera* cam1,
                                 const orbx_camera* cam2, const float* R1w, const float* t1w, const float* R2w, const float* t2w,
                                 const float* sigma2, const float* scaleFactors, int nlevels, int onlyStereo, int coarse,
                                 int checkOri, int32_t* match12, int32_t* nmatches) {
  KeyFrame K1, K2;
  fill_keyframe(K1, K1d, scaleFactors, nlevels);
  fill_keyframe(K2, K2d, scaleFactors, nlevels);
  K1.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
  K2.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
  GeometricCamera c1(cam1->fx, cam1->fy, cam1->cx, cam1->cy), c2(cam2->fx, cam2->fy, cam2->cx, cam2->cy);
  K1.mpCamera = &c1; K2.mpCamera = &c2;
  K1.fx = cam1->fx; K1.fy = cam1->fy; K1.cx = cam1->cx; K1.cy = cam1->cy; K1.mbf = cam1->bf; K1.mb = cam1->b;
  K2.fx = cam2->fx; K2.fy = cam2->fy; K2.cx = cam2->cx; K2.cy = cam2->cy; K2.mbf = cam2->bf; K2.mb = cam2->b;
  auto pose = [](KeyFrame& K, const float* R, const float* t) {
    K.Tcw = cv::Mat::eye(4, 4, CV_32F);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) K.Tcw.at<float>(i, j) = R[3 * i + j]; K.Tcw.at<float>(i, 3) = t[i]; }
    cv::Mat Rwc = K.Tcw.rowRange(0, 3).colRange(0, 3).t();       // KeyFrame::SetPose (src/KeyFrame.cc:108-116): Ow = -Rwc*tcw
    K.Ow = -Rwc * K.Tcw.rowRange(0, 3).col(3);
  };
  pose(K1, R1w, t1w);
  pose(K2, R2w, t2w);
  MapPoint dummy;
  for (int i = 0; i < K1d->n; ++i) if (has1[i]) K1.mvpMapPoints[i] = &dummy;
  for (int i = 0; i < K2d->n; ++i) if (has2[i]) K2.mvpMapPoints[i] = &dummy;
  fill_featvec(K1.mFeatVec, nn1, n1id, n1off, n1idx);
  fill_featvec(K2.mFeatVec, nn2, n2id, n2off, n2idx);
  std::vector<std::pair<size_t, size_t>> pairs;
  ORBmatcher matcher(0.6f, checkOri != 0);
  *nmatches = matcher.SearchForTriangulation(&K1, &K2, cv::Mat(), pairs, onlyStereo != 0, coarse != 0);
  for (int i = 0; i < K1d->n; ++i) match12[i] = -1;
  for (auto& pr : pairs) match12[pr.first] = (int)pr.second;
  return 0;
}

// src/ORBmatcher.cc:323-591
int ref_search_by_bow(const orbx_frame_desc* KFd, const orbx_frame_desc* Fd, const uint8_t* kf_has_mp, int nnK, const int32_t* fvK_node,
                      const int32_t* fvK_off, const int32_t* fvK_idx, int nnF, const int32_t* fvF_node, const int32_t* fvF_off,
                      const int32_t* fvF_idx, float nnratio, int check_orientation, int32_t* match_f, int32_t* nmatches) {
  KeyFrame K;
  Frame F;
  const float one[1] = {1.f};
  fill_keyframe(K, KFd, one, 1);
  fill_frame(F, Fd, one, 1);
  std::vector<MapPoint> mps(KFd->n);
  for (int i = 0; i < KFd->n; ++i) if (kf_has_mp[i]) K.mvpMapPoints[i] = &mps[i];
  fill_featvec(K.mFeatVec, nnK, fvK_node, fvK_off, fvK_idx);
  fill_featvec(F.mFeatVec, nnF, fvF_node, fvF_off, fvF_idx);
  std::vector<MapPoint*> out;
  ORBmatcher matcher(nnratio, check_orientation != 0);
  *nmatches = matcher.SearchByBoW(&K, F, out);
  for (int j = 0; j < Fd->n; ++j) match_f[j] = out[j] ? (int)(out[j] - mps.data()) : -1;
  return 0;
}

// src/ORBmatcher.cc:1630-1883 (bRight = false)
int ref_fuse(const orbx_frame_desc* KFd, const orbx_camera* cam, const float* Rcw, const float* tcw, const float* Ow, int nmp,
             const uint8_t* flags, const float* xw, const float* mp_max_dist, const float* mp_min_dist, const float* mp_normal,
             const uint8_t* mp_desc, float th, const float* scale_factors, const float* inv_level_sigma2, int nlevels,
             float log_scale_factor, int32_t* best_idx, int32_t* nfused) {
  KeyFrame K;
  fill_keyframe(K, KFd, scale_factors, nlevels);
  K.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + nlevels);
  K.mfLogScaleFactor = log_scale_factor;
  GeometricCamera camera(cam->fx, cam->fy, cam->cx, cam->cy);
  K.mpCamera = &camera;
  K.fx = cam->fx; K.fy = cam->fy; K.cx = cam->cx; K.cy = cam->cy; K.mbf = cam->bf; K.mb = cam->b;
  K.Tcw = cv::Mat::eye(4, 4, CV_32F);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) K.Tcw.at<float>(i, j) = Rcw[3 * i + j]; K.Tcw.at<float>(i, 3) = tcw[i]; }
  K.Ow = mat3x1(Ow);
  std::vector<MapPoint> mps(nmp);
  std::vector<MapPoint*> vp(nmp, nullptr);
  KeyFrame other;   // an observation in some other keyframe: IsInKeyFrame(pKF) is false
  for (int i = 0; i < nmp; ++i) {
    if (!(flags[i] & 1)) continue;   // NULL / bad / already in pKF: the reference skips all three the same way
    MapPoint& m = mps[i];
    m.mWorldPos = mat3x1(xw + 3 * i);
    m.mNormalVector = mat3x1(mp_normal + 3 * i);
    m.mDescriptor = desc_row(mp_desc + (size_t)32 * i);
    m.mfMaxDistance = mp_max_dist[i];
    m.mfMinDistance = mp_min_dist[i];
    m.nObs = 1;
    vp[i] = &m;
  }
  ORBmatcher matcher(0.6f, true);
  *nfused = matcher.Fuse(&K, vp, th, false);
  for (int i = 0; i < nmp; ++i) best_idx[i] = (vp[i] && !mps[i].addedObs.empty()) ? mps[i].addedObs[0].second : -1;
  return 0;
}

}  // extern "C"
