// Optimizer_orbx.cc — drop-in replacements for Optimizer::PoseOptimization (src/Optimizer.cc:907-1272),
// Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1811-2523) and Optimizer::PoseInertialOptimizationLastKeyFrame /
// LastFrame (src/Optimizer.cc:7665-8066, :8068-8603); the rest of the class stays in the reference.
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

// ---- visual-inertial pose refinement (pinhole rigs, Nleft == -1) ----------------------------------------------------
// Everything that the reference's vertex / edge constructors derive from the Frame, the KeyFrame and the pre-integration
// (ImuCamPose, delta measurements, the eigenvalue-clamped information matrices) is obtained from those very
// constructors; the Gauss-Newton rounds, the chi2 classification and the Hessian / marginalisation are delegated.
namespace {
struct VisualEdges {
  std::vector<float> xw, obs, isg;
  std::vector<uint8_t> closePt;
  std::vector<int> index;
};
void collect_visual_edges(Frame* pFrame, VisualEdges& V) {   // src/Optimizer.cc:7719-7823 == :8161-8273
  if (pFrame->Nleft != -1) throw std::runtime_error("orbx: PoseInertialOptimization* covers pinhole rigs (Nleft == -1)");
  const int N = pFrame->N;
  std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);
  for (int i = 0; i < N; i++) {
    MapPoint* pMP = pFrame->mvpMapPoints[i];
    if (!pMP) continue;
    pFrame->mvbOutlier[i] = false;
    const cv::KeyPoint& kpUn = pFrame->mvKeysUn[i];
    Eigen::Vector2d o2;
    o2(0) = kpUn.pt.x;
    o2(1) = kpUn.pt.y;
    const float unc2 = pFrame->mpCamera->uncertainty2(o2);    // 1 for Pinhole
    const cv::Mat Xw = pMP->GetWorldPos();
    for (int k = 0; k < 3; ++k) V.xw.push_back(Xw.at<float>(k));
    V.obs.push_back(kpUn.pt.x);
    V.obs.push_back(kpUn.pt.y);
    V.obs.push_back(pFrame->mvuRight[i]);                        // < 0: EdgeMonoOnlyPose, else EdgeStereoOnlyPose
    V.isg.push_back(pFrame->mvInvLevelSigma2[kpUn.octave] / unc2);
    V.closePt.push_back(pMP->mTrackDepth < 10.f ? 1 : 0);        // :7912 / :8412
    V.index.push_back(i);
  }
}
void body_state(const ImuCamPose& P, const cv::Mat& vel, const IMU::Bias& b, double* s) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) s[r * 3 + c] = P.Rwb(r, c);
    s[9 + r] = P.twb(r);
    s[12 + r] = vel.at<float>(r);
  }
  s[15] = b.bwx; s[16] = b.bwy; s[17] = b.bwz; s[18] = b.bax; s[19] = b.bay; s[20] = b.baz;
}
void information_matrices(IMU::Preintegrated* pInt, double* infoI, double* infoG, double* infoA) {
  const EdgeInertial ei(pInt);                                    // inverse + eigenvalue clamp (src/G2oTypes.cc:706-727)
  for (int r = 0; r < 9; ++r)
    for (int c = 0; c < 9; ++c) infoI[r * 9 + c] = ei.information()(r, c);
  const cv::Mat cvInfoG = pInt->C.rowRange(9, 12).colRange(9, 12).inv(cv::DECOMP_SVD);       // :7861 / :8316
  const cv::Mat cvInfoA = pInt->C.rowRange(12, 15).colRange(12, 15).inv(cv::DECOMP_SVD);     // :7872 / :8334
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { infoG[r * 3 + c] = cvInfoG.at<float>(r, c); infoA[r * 3 + c] = cvInfoA.at<float>(r, c); }
}
void mat3_to(const cv::Mat& M, double* out) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out[r * 3 + c] = M.at<float>(r, c);
}
// recover pose, velocity, biases and the prior for the next frame (:8022-8063 / :8533-8598)
void write_back(Frame* pFrame, const VisualEdges& V, const std::vector<uint8_t>& outlier, const double* state, const double* H15) {
  for (size_t e = 0; e < V.index.size(); ++e) pFrame->mvbOutlier[V.index[e]] = outlier[e] != 0;
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
}
}  // namespace

int Optimizer::PoseInertialOptimizationLastKeyFrame(Frame* pFrame, bool bRecInit) {
  VisualEdges V;
  collect_visual_edges(pFrame, V);
  const int E = (int)V.index.size();
  KeyFrame* pKF = pFrame->mpLastKeyFrame;
  const VertexPose VP(pFrame), VPk(pKF);                          // ImuCamPose(Frame*) / ImuCamPose(KeyFrame*)
  double state[21], kf[21], preint[16], infoI[81], infoG[9], infoA[9];
  body_state(VP.estimate(), pFrame->mVw, pFrame->mImuBias, state);
  const IMU::Bias bK = pKF->GetImuBias();
  body_state(VPk.estimate(), pKF->GetVelocity(), bK, kf);
  IMU::Preintegrated* pInt = pFrame->mpImuPreintegrated;
  const cv::Mat dR = pInt->GetDeltaRotation(bK), dV = pInt->GetDeltaVelocity(bK), dP = pInt->GetDeltaPosition(bK);   // :739-741
  mat3_to(dR, preint);
  for (int r = 0; r < 3; ++r) { preint[9 + r] = dV.at<float>(r); preint[12 + r] = dP.at<float>(r); }
  preint[15] = pInt->dT;
  information_matrices(pInt, infoI, infoG, infoA);
  float Tcw[16], Tcb[16], Tbc[16];
  pose_to_array(pFrame->mTcw, Tcw);
  pose_to_array(pFrame->mImuCalib.Tcb, Tcb);
  pose_to_array(pFrame->mImuCalib.Tbc, Tbc);
  orbx_camera cam{pFrame->fx, pFrame->fy, pFrame->cx, pFrame->cy, pFrame->mbf, pFrame->mb};
  std::vector<uint8_t> outlier(E > 0 ? E : 1, 0);
  double H15[225];
  int32_t nRet = 0, iters[4];
  orbx_shim::check("orbx_pose_inertial_optimization_last_keyframe",
                   orbx_pose_inertial_optimization_last_keyframe(orbx_shim::context(), E, V.xw.data(), V.obs.data(), V.isg.data(),
                                                                 V.closePt.data(), &cam, Tcw, Tcb, Tbc, state, kf, preint, infoI,
                                                                 infoG, infoA, bRecInit ? 1 : 0, outlier.data(), H15, &nRet,
                                                                 iters));
  write_back(pFrame, V, outlier, state, H15);
  return nRet;
}

int Optimizer::PoseInertialOptimizationLastFrame(Frame* pFrame, bool bRecInit) {
  VisualEdges V;
  collect_visual_edges(pFrame, V);
  const int E = (int)V.index.size();
  Frame* pFp = pFrame->mpPrevFrame;
  const VertexPose VP(pFrame), VPk(pFp);
  double state[21], prev[21], preint[16], jac[45], bias[6], infoI[81], infoG[9], infoA[9], prior[21], priorH[225];
  body_state(VP.estimate(), pFrame->mVw, pFrame->mImuBias, state);
  body_state(VPk.estimate(), pFp->mVw, pFp->mImuBias, prev);
  IMU::Preintegrated* pInt = pFrame->mpImuPreintegratedFrame;     // :8303
  mat3_to(pInt->dR, preint);
  for (int r = 0; r < 3; ++r) { preint[9 + r] = pInt->dV.at<float>(r); preint[12 + r] = pInt->dP.at<float>(r); }
  preint[15] = pInt->dT;
  mat3_to(pInt->JRg, jac); mat3_to(pInt->JVg, jac + 9); mat3_to(pInt->JVa, jac + 18); mat3_to(pInt->JPg, jac + 27); mat3_to(pInt->JPa, jac + 36);
  const IMU::Bias b0 = pInt->GetOriginalBias();                   // the bias the deltas were integrated with (`b`, src/ImuTypes.cc:367)
  bias[0] = b0.bwx; bias[1] = b0.bwy; bias[2] = b0.bwz; bias[3] = b0.bax; bias[4] = b0.bay; bias[5] = b0.baz;
  information_matrices(pInt, infoI, infoG, infoA);
  const ConstraintPoseImu* c = pFp->mpcpi;                        // EdgePriorPoseImu(pFp->mpcpi), :8345
  if (!c) throw std::runtime_error("orbx: PoseInertialOptimizationLastFrame needs pFrame->mpPrevFrame->mpcpi");
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) prior[r * 3 + k] = c->Rwb(r, k);
    prior[9 + r] = c->twb(r); prior[12 + r] = c->vwb(r); prior[15 + r] = c->bg(r); prior[18 + r] = c->ba(r);
  }
  for (int r = 0; r < 15; ++r)
    for (int k = 0; k < 15; ++k) priorH[r * 15 + k] = c->H(r, k);
  float Tcw[16], Tcb[16], Tbc[16];
  pose_to_array(pFrame->mTcw, Tcw);
  pose_to_array(pFrame->mImuCalib.Tcb, Tcb);
  pose_to_array(pFrame->mImuCalib.Tbc, Tbc);
  orbx_camera cam{pFrame->fx, pFrame->fy, pFrame->cx, pFrame->cy, pFrame->mbf, pFrame->mb};
  std::vector<uint8_t> outlier(E > 0 ? E : 1, 0);
  double H15[225];
  int32_t nRet = 0, iters[4];
  orbx_shim::check("orbx_pose_inertial_optimization_last_frame",
                   orbx_pose_inertial_optimization_last_frame(orbx_shim::context(), E, V.xw.data(), V.obs.data(), V.isg.data(),
                                                              V.closePt.data(), &cam, Tcw, Tcb, Tbc, state, prev, preint, jac, bias,
                                                              infoI, infoG, infoA, prior, priorH, bRecInit ? 1 : 0, outlier.data(),
                                                              H15, &nRet, iters));
  write_back(pFrame, V, outlier, state, H15);
  delete pFp->mpcpi;                                              // :8599-8600
  pFp->mpcpi = NULL;
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
  if (num_fixedKF < 2) {
    // Force two fixed keyframes exactly as the reference does (src/Optimizer.cc:1906-1945): ONE pass that keeps a
    // running lowest id and, in the else-branch only, a running "second lowest" -- a keyframe displaced from the lowest
    // slot is NOT demoted to second (ids visited as 4,3,5 fix {3,5}, not {3,4}).  The reference leaves both pointers
    // uninitialised when no candidate qualifies and pushes them anyway; here a slot that was never assigned is skipped.
    long unsigned int lowerId = pKF->mnId, secondLowerId = pKF->mnId;
    KeyFrame *pLowerKf = NULL, *pSecondLowerKF = NULL;
    for (KeyFrame* pKFi : lLocalKeyFrames) {
      if (pKFi == pKF || pKFi->mnId == pMap->GetInitKFid()) continue;
      if (pKFi->mnId < lowerId) {
        lowerId = pKFi->mnId;
        pLowerKf = pKFi;
      } else if (pKFi->mnId < secondLowerId) {
        secondLowerId = pKFi->mnId;
        pSecondLowerKF = pKFi;
      }
    }
    if (pLowerKf) {
      lFixedCameras.push_back(pLowerKf);
      lLocalKeyFrames.remove(pLowerKf);
      num_fixedKF++;
    }
    if (num_fixedKF < 2 && pSecondLowerKF) {
      lFixedCameras.push_back(pSecondLowerKF);
      lLocalKeyFrames.remove(pSecondLowerKF);
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
