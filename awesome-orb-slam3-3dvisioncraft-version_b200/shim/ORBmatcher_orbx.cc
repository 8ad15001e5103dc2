// ORBmatcher_orbx.cc — drop-in replacements for the hot ORBmatcher members and the two callers
// either side of the path (§8 f2: SearchByBoW(KeyFrame*, Frame&), Fuse(KeyFrame*, vector<MapPoint*>, th, bRight)); the
// other members of the class (SearchBySim3, the Sim3 Fuse / SearchByProjection overloads, ...) stay in the reference's
// src/ORBmatcher.cc, from which exactly these bodies are removed.  Each function flattens the pointer graph once into the SoA the C
// ABI takes — under the locks the reference's getters take — and scatters the index results back.
#include "orbx_shim_config.h"
#include <cstring>

namespace ORB_SLAM3 {

namespace {

struct FlatFrame {
  std::vector<orbx_keypoint> kps;
  std::vector<float> uright;
  orbx_frame_desc d;
};

template <typename FrameLike>
void flatten(const FrameLike& F, const std::vector<cv::KeyPoint>& keysUn, const cv::Mat& descriptors,
             const std::vector<float>& uRight, FlatFrame& out) {
  const int n = (int)keysUn.size();
  out.kps.resize(n);
  for (int i = 0; i < n; ++i) {
    const cv::KeyPoint& k = keysUn[i];
    out.kps[i] = orbx_keypoint{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave};
  }
  out.uright.assign(uRight.begin(), uRight.end());
  out.d.n = n;
  out.d.kps = out.kps.data();
  out.d.desc = descriptors.data;          // N x 32 CV_8U, continuous (ORBextractor creates it that way)
  out.d.uright = out.uright.empty() ? nullptr : out.uright.data();
  out.d.min_x = F.mnMinX;
  out.d.min_y = F.mnMinY;
  out.d.max_x = F.mnMaxX;
  out.d.max_y = F.mnMaxY;
}

void pose_to_array(const cv::Mat& T, float out[16]) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[r * 4 + c] = T.at<float>(r, c);
}

}  // namespace

// src/ORBmatcher.cc:2700-2716
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
  return orbx_descriptor_distance(a.ptr<uint8_t>(), b.ptr<uint8_t>());
}

// src/ORBmatcher.cc:59-255 (pinhole, Nleft == -1)
int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th,
                                   const bool bFarPoints, const float thFarPoints) {
  FlatFrame ff;
  flatten(F, F.mvKeysUn, F.mDescriptors, F.mvuRight, ff);
  const int n = ff.d.n, nq = (int)vpMapPoints.size();
  std::vector<uint8_t> blocked(n, 0), flags(nq, 0), desc((size_t)nq * 32);
  for (int i = 0; i < n; ++i)
    if (F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0) blocked[i] = 1;
  std::vector<float> px(nq), py(nq), pxr(nq), vc(nq);
  std::vector<int32_t> lvl(nq, 0), best(nq, -1);
  for (int q = 0; q < nq; ++q) {
    MapPoint* pMP = vpMapPoints[q];
    // the MapPoint's track fields are passed verbatim, never recomputed
    if (!pMP->mbTrackInView) continue;
    if (bFarPoints && pMP->mTrackDepth > thFarPoints) continue;
    if (pMP->isBad()) continue;
    flags[q] = 1 | (pMP->Observations() > 0 ? 2 : 0);
    px[q] = pMP->mTrackProjX;
    py[q] = pMP->mTrackProjY;
    pxr[q] = pMP->mTrackProjXR;
    lvl[q] = pMP->mnTrackScaleLevel;
    vc[q] = pMP->mTrackViewCos;
    const cv::Mat d = pMP->GetDescriptor();
    std::memcpy(&desc[(size_t)q * 32], d.data, 32);
  }
  int32_t nmatches = 0;
  orbx_shim::check("orbx_search_by_projection_map",
                   orbx_search_by_projection_map(orbx_shim::context(), &ff.d, blocked.data(), nq, px.data(), py.data(),
                                                 pxr.data(), lvl.data(), vc.data(), desc.data(), flags.data(), th,
                                                 mfNNratio, F.mvScaleFactors.data(), (int)F.mvScaleFactors.size(),
                                                 best.data(), &nmatches));
  for (int q = 0; q < nq; ++q)
    if (best[q] >= 0) F.mvpMapPoints[best[q]] = vpMapPoints[q];
  return nmatches;
}

// src/ORBmatcher.cc:2244-2509 (pinhole, Nleft == -1)
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
  FlatFrame ff;
  flatten(CurrentFrame, CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors, CurrentFrame.mvuRight, ff);
  const int n = ff.d.n, nq = LastFrame.N;
  std::vector<uint8_t> blocked(n, 0), flags(nq, 0), desc((size_t)nq * 32), kept(nq, 0);
  for (int i = 0; i < n; ++i)
    if (CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0) blocked[i] = 1;
  std::vector<float> xw((size_t)nq * 3), ang(nq);
  std::vector<int32_t> oct(nq, 0), match(nq, -1), curMatch(n, -1);
  for (int q = 0; q < nq; ++q) {
    MapPoint* pMP = LastFrame.mvpMapPoints[q];
    if (!pMP || LastFrame.mvbOutlier[q]) continue;
    flags[q] = 1 | (pMP->Observations() > 0 ? 2 : 0);
    const cv::Mat x3Dw = pMP->GetWorldPos();
    for (int k = 0; k < 3; ++k) xw[(size_t)q * 3 + k] = x3Dw.at<float>(k);
    oct[q] = LastFrame.mvKeysUn[q].octave;
    ang[q] = LastFrame.mvKeysUn[q].angle;
    const cv::Mat d = pMP->GetDescriptor();
    std::memcpy(&desc[(size_t)q * 32], d.data, 32);
  }
  float Tc[16], Tl[16];
  pose_to_array(CurrentFrame.mTcw, Tc);
  pose_to_array(LastFrame.mTcw, Tl);
  orbx_camera cam{CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy, CurrentFrame.mbf, CurrentFrame.mb};
  int32_t nmatches = 0;
  orbx_shim::check("orbx_search_by_projection_frame",
                   orbx_search_by_projection_frame(orbx_shim::context(), &ff.d, blocked.data(), &cam, Tc, Tl, nq,
                                                   flags.data(), xw.data(), oct.data(), ang.data(), desc.data(), th,
                                                   bMono ? 1 : 0, mbCheckOrientation ? 1 : 0,
                                                   CurrentFrame.mvScaleFactors.data(),
                                                   (int)CurrentFrame.mvScaleFactors.size(), match.data(), kept.data(),
                                                   curMatch.data(), &nmatches));
  // replay: assignments in increasing q, then the rotation filter's removals (:2379-2383, :2486-2505)
  for (int q = 0; q < nq; ++q)
    if (match[q] >= 0) CurrentFrame.mvpMapPoints[match[q]] = LastFrame.mvpMapPoints[q];
  for (int q = 0; q < nq; ++q)
    if (match[q] >= 0 && !kept[q]) CurrentFrame.mvpMapPoints[match[q]] = static_cast<MapPoint*>(NULL);
  return nmatches;
}

// src/ORBmatcher.cc:1138-1428 (pinhole keyframes, mpCamera2 == NULL)
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat /*F12: unused by the reference too*/,
                                       std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo,
                                       const bool bCoarse) {
  FlatFrame f1, f2;
  flatten(*pKF1, pKF1->mvKeysUn, pKF1->mDescriptors, pKF1->mvuRight, f1);
  flatten(*pKF2, pKF2->mvKeysUn, pKF2->mDescriptors, pKF2->mvuRight, f2);
  std::vector<uint8_t> has1(f1.d.n), has2(f2.d.n);
  for (int i = 0; i < f1.d.n; ++i) has1[i] = pKF1->GetMapPoint(i) != NULL;
  for (int i = 0; i < f2.d.n; ++i) has2[i] = pKF2->GetMapPoint(i) != NULL;
  // DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>) -> CSR, ascending node id
  auto to_csr = [](const DBoW2::FeatureVector& fv, std::vector<int32_t>& id, std::vector<int32_t>& off,
                   std::vector<int32_t>& idx) {
    off.push_back(0);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
      id.push_back((int32_t)it->first);
      for (size_t k = 0; k < it->second.size(); ++k) idx.push_back((int32_t)it->second[k]);
      off.push_back((int32_t)idx.size());
    }
  };
  std::vector<int32_t> id1, off1, idx1, id2, off2, idx2;
  to_csr(pKF1->mFeatVec, id1, off1, idx1);
  to_csr(pKF2->mFeatVec, id2, off2, idx2);
  float R1[9], t1[3], R2[9], t2[3];
  const cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) { R1[r * 3 + c] = R1w.at<float>(r, c); R2[r * 3 + c] = R2w.at<float>(r, c); }
    t1[r] = t1w.at<float>(r);
    t2[r] = t2w.at<float>(r);
  }
  orbx_camera c1{pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy, pKF1->mbf, pKF1->mb};
  orbx_camera c2{pKF2->fx, pKF2->fy, pKF2->cx, pKF2->cy, pKF2->mbf, pKF2->mb};
  std::vector<int32_t> m12(f1.d.n, -1);
  int32_t nmatches = 0;
  orbx_shim::check("orbx_search_for_triangulation",
                   orbx_search_for_triangulation(orbx_shim::context(), &f1.d, &f2.d, has1.data(), has2.data(),
                                                 (int)id1.size(), id1.data(), off1.data(), idx1.data(), (int)id2.size(),
                                                 id2.data(), off2.data(), idx2.data(), &c1, &c2, R1, t1, R2, t2,
                                                 pKF2->mvLevelSigma2.data(), pKF2->mvScaleFactors.data(),
                                                 (int)pKF2->mvScaleFactors.size(), bOnlyStereo, bCoarse,
                                                 mbCheckOrientation, m12.data(), &nmatches));
  vMatchedPairs.clear();
  vMatchedPairs.reserve(nmatches);
  for (size_t i = 0; i < m12.size(); ++i)
    if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));
  return nmatches;
}

namespace {
// DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>) -> CSR, ascending node id
void fv_to_csr(const DBoW2::FeatureVector& fv, std::vector<int32_t>& id, std::vector<int32_t>& off, std::vector<int32_t>& idx) {
  off.push_back(0);
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
    id.push_back((int32_t)it->first);
    for (size_t k = 0; k < it->second.size(); ++k) idx.push_back((int32_t)it->second[k]);
    off.push_back((int32_t)idx.size());
  }
}
}  // namespace

// src/ORBmatcher.cc:323-591 (pinhole, F.Nleft == -1)
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
  const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
  vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
  FlatFrame fk, ff;
  flatten(*pKF, pKF->mvKeysUn, pKF->mDescriptors, pKF->mvuRight, fk);
  flatten(F, F.mvKeys, F.mDescriptors, F.mvuRight, ff);          // the rotation check reads F.mvKeys (:497)
  std::vector<uint8_t> has(fk.d.n, 0);
  for (int i = 0; i < fk.d.n; ++i) has[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
  std::vector<int32_t> idK, offK, idxK, idF, offF, idxF, match(F.N > 0 ? F.N : 1, -1);
  fv_to_csr(pKF->mFeatVec, idK, offK, idxK);
  fv_to_csr(F.mFeatVec, idF, offF, idxF);
  int32_t nmatches = 0;
  orbx_shim::check("orbx_search_by_bow",
                   orbx_search_by_bow(orbx_shim::context(), &fk.d, &ff.d, has.data(), (int)idK.size(), idK.data(), offK.data(),
                                      idxK.data(), (int)idF.size(), idF.data(), offF.data(), idxF.data(), mfNNratio,
                                      mbCheckOrientation ? 1 : 0, match.data(), &nmatches));
  for (int j = 0; j < F.N; ++j)
    if (match[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[match[j]];
  return nmatches;
}

// src/ORBmatcher.cc:1630-1883 (bRight == false; the stereo-fisheye right-camera variant stays in the reference)
int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th, const bool bRight) {
  assert(!bRight);
  FlatFrame fk;
  flatten(*pKF, pKF->mvKeysUn, pKF->mDescriptors, pKF->mvuRight, fk);
  const int nMPs = (int)vpMapPoints.size();
  std::vector<uint8_t> flags(nMPs, 0), desc((size_t)nMPs * 32);
  std::vector<float> xw((size_t)nMPs * 3), nrm((size_t)nMPs * 3), maxd(nMPs), mind(nMPs);
  for (int i = 0; i < nMPs; ++i) {
    MapPoint* pMP = vpMapPoints[i];
    if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
    flags[i] = 1;
    const cv::Mat p = pMP->GetWorldPos(), n = pMP->GetNormal(), d = pMP->GetDescriptor();
    for (int k = 0; k < 3; ++k) { xw[(size_t)i * 3 + k] = p.at<float>(k); nrm[(size_t)i * 3 + k] = n.at<float>(k); }
    maxd[i] = pMP->GetMaxDistance();
    mind[i] = pMP->GetMinDistance();
    std::memcpy(&desc[(size_t)i * 32], d.data, 32);
  }
  float R[9], t[3], O[3];
  const cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = Rcw.at<float>(r, c);
    t[r] = tcw.at<float>(r);
    O[r] = Ow.at<float>(r);
  }
  orbx_camera cam{pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mbf, pKF->mb};
  std::vector<int32_t> best(nMPs > 0 ? nMPs : 1, -1);
  int32_t nHits = 0;
  orbx_shim::check("orbx_fuse",
                   orbx_fuse(orbx_shim::context(), &fk.d, &cam, R, t, O, nMPs, flags.data(), xw.data(), maxd.data(), mind.data(),
                             nrm.data(), desc.data(), th, pKF->mvScaleFactors.data(), pKF->mvInvLevelSigma2.data(),
                             (int)pKF->mvScaleFactors.size(), pKF->mfLogScaleFactor, best.data(), &nHits));
  // the map surgery of :1844-1867, replayed in the reference's order over the search results
  int nFused = 0;
  for (int i = 0; i < nMPs; ++i) {
    if (best[i] < 0) continue;
    MapPoint* pMP = vpMapPoints[i];
    if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;   // an earlier Replace() may have retired it
    MapPoint* pMPinKF = pKF->GetMapPoint(best[i]);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) {
        if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
        else pMPinKF->Replace(pMP);
      }
    } else {
      pMP->AddObservation(pKF, best[i]);
      pKF->AddMapPoint(pMP, best[i]);
    }
    ++nFused;
  }
  return nFused;
}

}  // namespace ORB_SLAM3
