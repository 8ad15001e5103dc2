// Optimizer_orbx.cc — drop-in replacements for Optimizer::PoseOptimization (src/Optimizer.cc:907-1272),
// Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1811-2523) and Optimizer::PoseInertialOptimizationLastKeyFrame
// (src/Optimizer.cc:7665-8066); the rest of the class stays in the reference.
// Graph *collection* and *write-back* walk the reference's pointer graph exactly as the reference does (same
// locks, same bookkeeping fields); only the numeric core — g2o graph, LM, Schur, chi2 tests — is delegated.
#include "orbx_shim_config.h"
#include <list>
#include <map>
#include <set>
#include <stdexcept>

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

// Visual-inertial pose refinement against the last keyframe (pinhole rigs, Nleft == -1).  Everything that the reference's
// vertex / edge constructors derive from the Frame, the KeyFrame and the pre-integration (ImuCamPose, the delta
// measurements at the keyframe's bias, the eigenvalue-clamped information matrices) is obtained from those very
// constructors; the Gauss-Newton rounds, the chi2 classification and the 15x15 Hessian are delegated.
int Optimizer::PoseInertialOptimizationLastKeyFrame(Frame* pFrame, bool bRecInit) {
  const int N = pFrame->N;
  if (pFrame->Nleft != -1) throw std::runtime_error("orbx: PoseInertialOptimizationLastKeyFrame covers pinhole rigs (Nleft == -1)");
  std::vector<float> xw, obs, isg;
  std::vector<uint8_t> closePt;
  std::vector<int> index;
  {
    std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);   // src/Optimizer.cc:7720
    for (int i = 0; i < N; i++) {
      MapPoint* pMP = pFrame->mvpMapPoints[i];
      if (!pMP) continue;
      pFrame->mvbOutlier[i] = false;
      const cv::KeyPoint& kpUn = pFrame->mvKeysUn[i];
      Eigen::Vector2d o2;
      o2(0) = kpUn.pt.x;
      o2(1) = kpUn.pt.y;
      const float unc2 = pFrame->mpCamera->uncertainty2(o2);    // 1 for Pinhole (:7749, :7781)
      const cv::Mat Xw = pMP->GetWorldPos();
      for (int k = 0; k < 3; ++k) xw.push_back(Xw.at<float>(k));
      obs.push_back(kpUn.pt.x);
      obs.push_back(kpUn.pt.y);
      obs.push_back(pFrame->mvuRight[i]);                        // < 0: EdgeMonoOnlyPose, else EdgeStereoOnlyPose
      isg.push_back(pFrame->mvInvLevelSigma2[kpUn.octave] / unc2);
      closePt.push_back(pMP->mTrackDepth < 10.f ? 1 : 0);        // :7912
      index.push_back(i);
    }
  }
  const int E = (int)index.size();
  KeyFrame* pKF = pFrame->mpLastKeyFrame;
  const VertexPose VP(pFrame), VPk(pKF);                          // ImuCamPose(Frame*) / ImuCamPose(KeyFrame*)
  const ImuCamPose& P = VP.estimate();
  const ImuCamPose& Pk = VPk.estimate();
  double state[21], kf[21], preint[16], infoI[81], infoG[9], infoA[9];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) { state[r * 3 + c] = P.Rwb(r, c); kf[r * 3 + c] = Pk.Rwb(r, c); }
    state[9 + r] = P.twb(r);
    kf[9 + r] = Pk.twb(r);
    state[12 + r] = pFrame->mVw.at<float>(r);                    // VertexVelocity(Frame*) (src/G2oTypes.cc:668-672)
    kf[12 + r] = pKF->GetVelocity().at<float>(r);
  }
  const IMU::Bias bF = pFrame->mImuBias, bK = pKF->GetImuBias();
  const float gF[3] = {bF.bwx, bF.bwy, bF.bwz}, aF[3] = {bF.bax, bF.bay, bF.baz};
  const float gK[3] = {bK.bwx, bK.bwy, bK.bwz}, aK[3] = {bK.bax, bK.bay, bK.baz};
  for (int r = 0; r < 3; ++r) { state[15 + r] = gF[r]; state[18 + r] = aF[r]; kf[15 + r] = gK[r]; kf[18 + r] = aK[r]; }
  IMU::Preintegrated* pInt = pFrame->mpImuPreintegrated;
  const cv::Mat dR = pInt->GetDeltaRotation(bK), dV = pInt->GetDeltaVelocity(bK), dP = pInt->GetDeltaPosition(bK);   // :739-741
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) preint[r * 3 + c] = dR.at<float>(r, c);
    preint[9 + r] = dV.at<float>(r);
    preint[12 + r] = dP.at<float>(r);
  }
  preint[15] = pInt->dT;
  const EdgeInertial ei(pInt);                                    // inverse + eigenvalue clamp (src/G2oTypes.cc:706-727)
  for (int r = 0; r < 9; ++r)
    for (int c = 0; c < 9; ++c) infoI[r * 9 + c] = ei.information()(r, c);
  const cv::Mat cvInfoG = pInt->C.rowRange(9, 12).colRange(9, 12).inv(cv::DECOMP_SVD);       // :7861
  const cv::Mat cvInfoA = pInt->C.rowRange(12, 15).colRange(12, 15).inv(cv::DECOMP_SVD);     // :7872
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { infoG[r * 3 + c] = cvInfoG.at<float>(r, c); infoA[r * 3 + c] = cvInfoA.at<float>(r, c); }
  float Tcw[16], Tcb[16], Tbc[16];
  pose_to_array(pFrame->mTcw, Tcw);
  pose_to_array(pFrame->mImuCalib.Tcb, Tcb);
  pose_to_array(pFrame->mImuCalib.Tbc, Tbc);
  orbx_camera cam{pFrame->fx, pFrame->fy, pFrame->cx, pFrame->cy, pFrame->mbf, pFrame->mb};
  std::vector<uint8_t> outlier(E > 0 ? E : 1, 0);
  double H15[225];
  int32_t nRet = 0, iters[4];
  orbx_shim::check("orbx_pose_inertial_optimization_last_keyframe",
                   orbx_pose_inertial_optimization_last_keyframe(orbx_shim::context(), E, xw.data(), obs.data(), isg.data(),
                                                                 closePt.data(), &cam, Tcw, Tcb, Tbc, state, kf, preint, infoI,
                                                                 infoG, infoA, bRecInit ? 1 : 0, outlier.data(), H15, &nRet,
                                                                 iters));
  for (int e = 0; e < E; ++e) pFrame->mvbOutlier[index[e]] = outlier[e] != 0;
  // recover pose, velocity, biases and the prior for the next frame (:8022-8063)
  Eigen::Matrix3d Rwb;
  Eigen::Vector3d twb, vwb, bg, ba;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Rwb(r, c) = state[r * 3 + c];
    twb(r) = state[9 + r]; vwb(r) = state[12 + r]; bg(r) = state[15 + r]; ba(r) = state[18 + r];
  }
  pFrame->SetImuPoseVelocity(Converter::toCvMat(Rwb), Converter::toCvMat(twb), Converter::toCvMat(vwb));
  pFrame->mImuBias = IMU::Bias(state[18], state[19], state[20], state[15], state[16], state[17]);
  Matrix15d H;
  for (int r = 0; r < 15; ++r)
    for (int c = 0; c < 15; ++c) H(r, c) = H15[r * 15 + c];
  pFrame->mpcpi = new ConstraintPoseImu(Rwb, twb, vwb, bg, ba, H);
  return nRet;
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
