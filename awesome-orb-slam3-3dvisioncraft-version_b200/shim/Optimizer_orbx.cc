// Optimizer_orbx.cc — drop-in replacements for Optimizer::PoseOptimization (src/Optimizer.cc:907-1272) and
// Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1811-2523); the rest of the class stays in the reference.
// Graph *collection* and *write-back* walk the reference's pointer graph exactly as the reference does (same
// locks, same bookkeeping fields); only the numeric core — g2o graph, LM, Schur, chi2 tests — is delegated.
#include "orbx_shim_config.h"
#include <list>
#include <map>
#include <set>

namespace ORB_SLAM3 {

namespace {
void pose_to_array(const cv::Mat& T, float* out) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[r * 4 + c] = T.at<float>(r, c);
}
cv::Mat array_to_pose(const float* a) {
  cv::Mat T(4, 4, CV_32F);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) T.at<float>(r, c) = a[r * 4 + c];
  return T;
}
}  // namespace

int Optimizer::PoseOptimization(Frame* pFrame) {
  const int N = pFrame->N;
  std::vector<float> xw, obs, isg;
  std::vector<int> index;
  xw.reserve(3 * N);
  obs.reserve(3 * N);
  {
    std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);   // src/Optimizer.cc:959
    for (int i = 0; i < N; i++) {
      MapPoint* pMP = pFrame->mvpMapPoints[i];
      if (!pMP) continue;
      pFrame->mvbOutlier[i] = false;
      const cv::KeyPoint& kpUn = pFrame->mvKeysUn[i];
      const cv::Mat Xw = pMP->GetWorldPos();
      for (int k = 0; k < 3; ++k) xw.push_back(Xw.at<float>(k));
      obs.push_back(kpUn.pt.x);
      obs.push_back(kpUn.pt.y);
      obs.push_back(pFrame->mvuRight[i]);     // < 0 selects the monocular edge (:972), else the stereo edge (:1016)
      isg.push_back(pFrame->mvInvLevelSigma2[kpUn.octave]);
      index.push_back(i);
    }
  }
  const int E = (int)index.size();
  if (E < 3) return 0;
  float T[16];
  pose_to_array(pFrame->mTcw, T);
  orbx_camera cam{pFrame->fx, pFrame->fy, pFrame->cx, pFrame->cy, pFrame->mbf, pFrame->mb};
  std::vector<uint8_t> outlier(E, 0);
  int32_t nInliers = 0, iters[4];
  orbx_shim::check("orbx_pose_optimization",
                   orbx_pose_optimization(orbx_shim::context(), E, xw.data(), obs.data(), isg.data(), &cam, T,
                                          outlier.data(), &nInliers, iters));
  for (int e = 0; e < E; ++e) pFrame->mvbOutlier[index[e]] = outlier[e] != 0;
  pFrame->SetPose(array_to_pose(T));
  return nInliers;
}

void Optimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, int& num_fixedKF) {
  // ---- local keyframes, local map points, fixed keyframes: src/Optimizer.cc:1816-1945 ----
  std::list<KeyFrame*> lLocalKeyFrames;
  lLocalKeyFrames.push_back(pKF);
  pKF->mnBALocalForKF = pKF->mnId;
  Map* pCurrentMap = pKF->GetMap();
  const std::vector<KeyFrame*> vNeighKFs = pKF->GetVectorCovisibleKeyFrames();
  for (size_t i = 0; i < vNeighKFs.size(); i++) {
    KeyFrame* pKFi = vNeighKFs[i];
    pKFi->mnBALocalForKF = pKF->mnId;
    if (!pKFi->isBad() && pKFi->GetMap() == pCurrentMap) lLocalKeyFrames.push_back(pKFi);
  }
  num_fixedKF = 0;
  std::list<MapPoint*> lLocalMapPoints;
  for (KeyFrame* pKFi : lLocalKeyFrames) {
    if (pKFi->mnId == pMap->GetInitKFid()) num_fixedKF = 1;
    for (MapPoint* pMP : pKFi->GetMapPointMatches())
      if (pMP && !pMP->isBad() && pMP->GetMap() == pCurrentMap && pMP->mnBALocalForKF != pKF->mnId) {
        lLocalMapPoints.push_back(pMP);
        pMP->mnBALocalForKF = pKF->mnId;
      }
  }
  std::list<KeyFrame*> lFixedCameras;
  for (MapPoint* pMP : lLocalMapPoints) {
    const std::map<KeyFrame*, std::tuple<int, int> > observations = pMP->GetObservations();
    for (const auto& ob : observations) {
      KeyFrame* pKFi = ob.first;
      if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) {
        pKFi->mnBAFixedForKF = pKF->mnId;
        if (!pKFi->isBad() && pKFi->GetMap() == pCurrentMap) lFixedCameras.push_back(pKFi);
      }
    }
  }
  num_fixedKF = (int)lFixedCameras.size() + num_fixedKF;
  if (num_fixedKF < 2) {   // force two fixed keyframes: the two lowest ids of the window (:1901-1945)
    for (int pass = 0; pass < 2 && num_fixedKF < 2; ++pass) {
      KeyFrame* lowest = NULL;
      for (KeyFrame* pKFi : lLocalKeyFrames) {
        if (pKFi == pKF || pKFi->mnId == pMap->GetInitKFid()) continue;
        if (!lowest || pKFi->mnId < lowest->mnId) lowest = pKFi;
      }
      if (!lowest) break;
      lFixedCameras.push_back(lowest);
      lLocalKeyFrames.remove(lowest);
      num_fixedKF++;
    }
  }
  if (pbStopFlag && *pbStopFlag) return;

  // ---- flatten: vertices and one edge per observation (src/Optimizer.cc:1982-2193) ----
  std::map<KeyFrame*, int> kfIndex;
  std::vector<KeyFrame*> kfs;
  std::vector<uint8_t> fixed;
  for (KeyFrame* pKFi : lLocalKeyFrames) {
    kfIndex[pKFi] = (int)kfs.size();
    kfs.push_back(pKFi);
    fixed.push_back(pKFi->mnId == pMap->GetInitKFid() ? 1 : 0);
  }
  for (KeyFrame* pKFi : lFixedCameras) {
    kfIndex[pKFi] = (int)kfs.size();
    kfs.push_back(pKFi);
    fixed.push_back(1);
  }
  std::vector<float> kfT(16 * kfs.size());
  for (size_t k = 0; k < kfs.size(); ++k) pose_to_array(kfs[k]->GetPose(), &kfT[16 * k]);
  std::vector<MapPoint*> mps(lLocalMapPoints.begin(), lLocalMapPoints.end());
  std::vector<float> xyz(3 * mps.size()), eobs, eisg;
  std::vector<int32_t> ekf, emp;
  for (size_t m = 0; m < mps.size(); ++m) {
    const cv::Mat P = mps[m]->GetWorldPos();
    for (int k = 0; k < 3; ++k) xyz[3 * m + k] = P.at<float>(k);
    const std::map<KeyFrame*, std::tuple<int, int> > observations = mps[m]->GetObservations();
    for (const auto& ob : observations) {
      KeyFrame* pKFi = ob.first;
      if (pKFi->isBad() || pKFi->GetMap() != pCurrentMap) continue;
      const int leftIndex = std::get<0>(ob.second);
      if (leftIndex == -1) continue;
      const std::map<KeyFrame*, int>::const_iterator it = kfIndex.find(pKFi);
      if (it == kfIndex.end()) continue;      // not a vertex: g2o would reject the edge as well
      const cv::KeyPoint& kpUn = pKFi->mvKeysUn[leftIndex];
      ekf.push_back(it->second);
      emp.push_back((int32_t)m);
      eobs.push_back(kpUn.pt.x);
      eobs.push_back(kpUn.pt.y);
      eobs.push_back(pKFi->mvuRight[leftIndex]);   // >= 0: EdgeStereoSE3ProjectXYZ, else EdgeSE3ProjectXYZ
      eisg.push_back(pKFi->mvInvLevelSigma2[kpUn.octave]);
    }
  }
  if (kfs.empty() || mps.empty() || ekf.empty()) return;
  orbx_camera cam{pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mbf, pKF->mb};
  std::vector<uint8_t> bad(ekf.size(), 0);
  int32_t iters[2] = {0, 0}, status = 0;
  orbx_shim::check("orbx_local_ba",
                   orbx_local_ba(orbx_shim::context(), (int)kfs.size(), kfT.data(), fixed.data(), (int)mps.size(),
                                 xyz.data(), (int)ekf.size(), ekf.data(), emp.data(), eobs.data(), eisg.data(), &cam,
                                 pMap->IsInertial() ? 100.0 : 0.0, reinterpret_cast<const volatile uint8_t*>(pbStopFlag),
                                 bad.data(), iters, &status));
  if (status != 0) return;   // stopped before optimising, or the >= 50 % outlier sanity check (:2348-2352)

  // ---- write-back under the map mutex: src/Optimizer.cc:2375-2510 ----
  std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
  for (size_t e = 0; e < bad.size(); ++e)
    if (bad[e] && !mps[emp[e]]->isBad()) {
      kfs[ekf[e]]->EraseMapPointMatch(mps[emp[e]]);
      mps[emp[e]]->EraseObservation(kfs[ekf[e]]);
    }
  for (size_t k = 0; k < lLocalKeyFrames.size(); ++k)   // local keyframes come first in `kfs`; fixed ones keep their pose
    kfs[k]->SetPose(array_to_pose(&kfT[16 * k]));
  for (size_t m = 0; m < mps.size(); ++m) {
    cv::Mat P(3, 1, CV_32F);
    for (int k = 0; k < 3; ++k) P.at<float>(k) = xyz[3 * m + k];
    mps[m]->SetWorldPos(P);
    mps[m]->UpdateNormalAndDepth();
  }
}

}  // namespace ORB_SLAM3
