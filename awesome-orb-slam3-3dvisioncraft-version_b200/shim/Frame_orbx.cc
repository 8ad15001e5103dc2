// Frame_orbx.cc — drop-in replacements for Frame::ComputeStereoMatches (src/Frame.cc:955-1133), Frame::UndistortKeyPoints
// (:874-924) and a batched form of Frame::isInFrustum (:571-650) for Tracking::SearchLocalPoints.  The pyramids
// the reference reads through mpORBextractorLeft/Right->mvImagePyramid stay on the device: the two extractor
// instances that just produced mvKeys / mvKeysRight (src/Frame.cc:111-114) still hold them.
#include "orbx_shim_config.h"

namespace orbx_shim { orbx_ext* extractor_handle(const void* self); }

namespace ORB_SLAM3 {

void Frame::ComputeStereoMatches() {
  mvuRight = std::vector<float>(N, -1.0f);
  mvDepth = std::vector<float>(N, -1.0f);
  orbx_ext* L = orbx_shim::extractor_handle(mpORBextractorLeft);
  orbx_ext* R = orbx_shim::extractor_handle(mpORBextractorRight);
  if (!L || !R || N == 0) return;
  auto flat = [](const std::vector<cv::KeyPoint>& v) {
    std::vector<orbx_keypoint> o(v.size());
    for (size_t i = 0; i < v.size(); ++i)
      o[i] = orbx_keypoint{v[i].pt.x, v[i].pt.y, v[i].size, v[i].angle, v[i].response, v[i].octave};
    return o;
  };
  const std::vector<orbx_keypoint> kl = flat(mvKeys), kr = flat(mvKeysRight);
  orbx_shim::check("orbx_stereo_match",
                   orbx_stereo_match(orbx_shim::context(), L, 0, R, 0, kl.data(), mDescriptors.data, (int)kl.size(),
                                     kr.data(), mDescriptorsRight.data, (int)kr.size(), mbf, mb, mvuRight.data(),
                                     mvDepth.data()));
}

// src/Frame.cc:874-924
void Frame::UndistortKeyPoints() {
  if (mDistCoef.at<float>(0) == 0.0) {
    mvKeysUn = mvKeys;
    return;
  }
  std::vector<float> xy(2 * (size_t)N), d;
  for (int i = 0; i < N; ++i) { xy[2 * i] = mvKeys[i].pt.x; xy[2 * i + 1] = mvKeys[i].pt.y; }
  for (int i = 0; i < mDistCoef.rows; ++i) d.push_back(mDistCoef.at<float>(i));
  orbx_camera cam{fx, fy, cx, cy, mbf, mb};
  orbx_shim::check("orbx_undistort_keypoints",
                   orbx_undistort_keypoints(orbx_shim::context(), xy.data(), N, &cam, d.data(), (int)d.size(), xy.data()));
  mvKeysUn.resize(N);
  for (int i = 0; i < N; ++i) {
    cv::KeyPoint kp = mvKeys[i];
    kp.pt.x = xy[2 * i];
    kp.pt.y = xy[2 * i + 1];
    mvKeysUn[i] = kp;
  }
}

// Batched Frame::isInFrustum (src/Frame.cc:571-650, Nleft == -1): the loop of Tracking::SearchLocalPoints
// (src/Tracking.cc:2878-2900) calls this once with its candidate MapPoints instead of isInFrustum per point
// (a ~10-line edit of that loop); returns how many are in view and sets the same MapPoint fields.
int Frame::isInFrustumBatch(const std::vector<MapPoint*>& vpMP, float viewingCosLimit) {
  const int n = (int)vpMP.size();
  std::vector<float> xw(3 * (size_t)n), nrm(3 * (size_t)n), maxd(n), mind(n), px(n), py(n), pxr(n), dep(n), vc(n);
  std::vector<int32_t> lvl(n);
  std::vector<uint8_t> in(n);
  for (int i = 0; i < n; ++i) {
    MapPoint* p = vpMP[i];
    const cv::Mat P = p->GetWorldPos(), Pn = p->GetNormal();
    for (int k = 0; k < 3; ++k) { xw[3 * i + k] = P.at<float>(k); nrm[3 * i + k] = Pn.at<float>(k); }
    maxd[i] = p->GetMaxDistance();
    mind[i] = p->GetMinDistance();
    pxr[i] = p->mTrackProjXR; dep[i] = p->mTrackDepth; lvl[i] = p->mnTrackScaleLevel; vc[i] = p->mTrackViewCos;   // stale values survive
  }
  float R[9], t[3], O[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = mRcw.at<float>(r, c);
    t[r] = mtcw.at<float>(r);
    O[r] = mOw.at<float>(r);
  }
  orbx_camera cam{fx, fy, cx, cy, mbf, mb};
  int32_t cnt = 0;
  orbx_shim::check("orbx_is_in_frustum",
                   orbx_is_in_frustum(orbx_shim::context(), &cam, R, t, O, mnMinX, mnMaxX, mnMinY, mnMaxY, viewingCosLimit,
                                      mnScaleLevels, mfLogScaleFactor, n, xw.data(), maxd.data(), mind.data(), nrm.data(), in.data(),
                                      px.data(), py.data(), pxr.data(), dep.data(), lvl.data(), vc.data(), &cnt));
  for (int i = 0; i < n; ++i) {
    MapPoint* p = vpMP[i];
    p->mbTrackInView = in[i] != 0;
    p->mTrackProjX = px[i];
    p->mTrackProjY = py[i];
    if (in[i]) { p->mTrackProjXR = pxr[i]; p->mTrackDepth = dep[i]; p->mnTrackScaleLevel = lvl[i]; p->mTrackViewCos = vc[i]; }
  }
  return cnt;
}

}  // namespace ORB_SLAM3
