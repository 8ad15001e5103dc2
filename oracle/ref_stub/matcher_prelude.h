// matcher_prelude.h — TEST INFRASTRUCTURE.  Force-included (-include) in front of the reference's UNMODIFIED
// src/ORBmatcher.cc so that it compiles into oracle/_ref/libref_matcher.so without Eigen / Boost / g2o / DBoW2's
// vocabulary / the rest of ORB-SLAM3.
//
// include/ORBmatcher.h pulls in MapPoint.h, KeyFrame.h and Frame.h, whose transitive includes (Eigen, Boost
// serialization, g2o, Atlas ...) are not in this image.  This file pre-defines their include guards and supplies
// stand-in classes with exactly the members ORBmatcher.cc touches, as plain data the test glue fills.  What is restated
// here is STATE and the few helper methods that live in other reference files:
//   Frame::GetFeaturesInArea / KeyFrame::GetFeaturesInArea   src/Frame.cc:755-850, src/KeyFrame.cc:642-700  -> oracle grid
//   KeyFrame::IsInImage                                       src/KeyFrame.cc:702-705
//   MapPoint::PredictScale / Get{Min,Max}DistanceInvariance   src/MapPoint.cc:566-600
//   Pinhole::project / epipolarConstrain                      src/CameraModels/Pinhole.cpp:31-50,155-177
// Every matching decision — the loops, gates, ratio tests, rotation histograms, ComputeThreeMaxima,
// DescriptorDistance — is the reference's own compiled code.
#ifndef ORK_MATCHER_PRELUDE_H_
#define ORK_MATCHER_PRELUDE_H_
#define MAPPOINT_H
#define KEYFRAME_H
#define FRAME_H
#include "cvstub.h"
#include <map>
#include <set>
#include <tuple>
#include <vector>
#include <mutex>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;   // the reference headers replaced here leak it (include/ORBmatcher.h:73 relies on that)

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace ORB_SLAM3 {

class KeyFrame;
class Frame;
class MapPoint;

// Pinhole camera (src/CameraModels/Pinhole.cpp).  Arithmetic in float like the reference (mvParameters is vector<float>).
class GeometricCamera {
 public:
  float fx = 0, fy = 0, cx = 0, cy = 0;
  GeometricCamera() {}
  GeometricCamera(float fx_, float fy_, float cx_, float cy_) : fx(fx_), fy(fy_), cx(cx_), cy(cy_) {}
  virtual ~GeometricCamera() {}
  virtual cv::Point2f project(const cv::Point3f& p3D) {            // Pinhole.cpp:31-34
    return cv::Point2f(fx * p3D.x / p3D.z + cx, fy * p3D.y / p3D.z + cy);
  }
  virtual cv::Point2f project(const cv::Mat& m3D) {                // Pinhole.cpp:36-41
    const float* p3D = m3D.ptr<float>();
    return project(cv::Point3f(p3D[0], p3D[1], p3D[2]));
  }
  virtual cv::Point3f unproject(const cv::Point2f& p2D) {          // Pinhole.cpp:60-63
    return cv::Point3f((p2D.x - cx) / fy * 0 + (p2D.x - cx) / fx, (p2D.y - cy) / fy, 1.f);
  }
  virtual cv::Mat toK() {                                          // Pinhole.cpp:148-153
    cv::Mat K = (cv::Mat_<float>(3, 3) << fx, 0.f, cx, 0.f, fy, cy, 0.f, 0.f, 1.f);
    return K;
  }
  // Pinhole.cpp:155-177
  virtual bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& R12,
                                 const cv::Mat& t12, const float sigmaLevel, const float unc) {
    cv::Mat t12x = SkewSymmetricMatrix(t12);
    cv::Mat K1 = this->toK();
    cv::Mat K2 = pCamera2->toK();
    cv::Mat F12 = K1.t().inv() * t12x * R12 * K2.inv();
    const float a = kp1.pt.x * F12.at<float>(0, 0) + kp1.pt.y * F12.at<float>(1, 0) + F12.at<float>(2, 0);
    const float b = kp1.pt.x * F12.at<float>(0, 1) + kp1.pt.y * F12.at<float>(1, 1) + F12.at<float>(2, 1);
    const float c = kp1.pt.x * F12.at<float>(0, 2) + kp1.pt.y * F12.at<float>(1, 2) + F12.at<float>(2, 2);
    const float num = a * kp2.pt.x + b * kp2.pt.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * unc;
  }
  virtual bool matchAndtriangulate(const cv::KeyPoint&, const cv::KeyPoint&, GeometricCamera*, cv::Mat&, cv::Mat&, const float,
                                   const float, cv::Mat&) {
    return false;   // KannalaBrandt8 only; Pinhole::matchAndtriangulate returns false (include/CameraModels/Pinhole.h)
  }
  static cv::Mat SkewSymmetricMatrix(const cv::Mat& v) {           // Pinhole.cpp:184-189
    return (cv::Mat_<float>(3, 3) << 0, -v.at<float>(2), v.at<float>(1), v.at<float>(2), 0, -v.at<float>(0), -v.at<float>(1),
            v.at<float>(0), 0);
  }
};

// the oracle's restatement of the 64x48 grid query (Frame.cc:755-850 / KeyFrame.cc:642-700), provided by the glue
struct RefGrid;
RefGrid* ref_grid_build(const std::vector<cv::KeyPoint>& keysUn, float minX, float minY, float maxX, float maxY);
void ref_grid_free(RefGrid*);
std::vector<size_t> ref_grid_query(const RefGrid*, const std::vector<cv::KeyPoint>& keysUn, float x, float y, float r, int minLevel,
                                   int maxLevel);

class MapPoint {
 public:
  // tracking scratch written by Frame::isInFrustum (include/MapPoint.h:133-146)
  float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
  bool mbTrackInView = false, mbTrackInViewR = false;
  int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
  float mTrackViewCos = 0, mTrackViewCosR = 0;
  long unsigned int mnId = 0;
  long unsigned int mnLastFrameSeen = 0;
  // state
  cv::Mat mWorldPos, mNormalVector, mDescriptor;
  float mfMinDistance = 0, mfMaxDistance = 0;
  bool mbBad = false;
  int nObs = 0;
  std::map<KeyFrame*, std::tuple<int, int>> mObservations;
  MapPoint* mpReplaced = nullptr;
  // log of the map surgery ORBmatcher::Fuse performs (replayed by the caller in the product)
  std::vector<std::pair<KeyFrame*, int>> addedObs;

  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
  cv::Mat GetNormal() { return mNormalVector.clone(); }
  cv::Mat GetDescriptor() { return mDescriptor.clone(); }
  bool isBad() { return mbBad; }
  int Observations() { return nObs; }
  float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }   // src/MapPoint.cc:566-570
  float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }   // src/MapPoint.cc:572-576
  int PredictScale(const float& currentDist, KeyFrame* pKF);          // src/MapPoint.cc:578-594
  int PredictScale(const float& currentDist, Frame* pF);              // src/MapPoint.cc:596-612
  bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
  std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* pKF) {
    auto it = mObservations.find(pKF);
    return it == mObservations.end() ? std::tuple<int, int>(-1, -1) : it->second;
  }
  void AddObservation(KeyFrame* pKF, int idx) { addedObs.push_back(std::make_pair(pKF, idx)); }
  void Replace(MapPoint* pMP) { mpReplaced = pMP; }
  MapPoint* GetReplaced() { return mpReplaced; }
};

class Frame {
 public:
  int N = 0, Nleft = -1, Nright = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
  std::vector<float> mvuRight, mvDepth;
  cv::Mat mDescriptors, mDescriptorsRight;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
  cv::Mat mTcw, mTrl, mTlr;
  float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0, mThDepth = 0;
  float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
  int mnScaleLevels = 8;
  float mfScaleFactor = 1.2f, mfLogScaleFactor = 0;
  std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
  long unsigned int mnId = 0;
  RefGrid* grid = nullptr;

  ~Frame() { if (grid) ref_grid_free(grid); }
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                        const int maxLevel = -1, const bool bRight = false) const {
    (void)bRight;   // Nleft == -1 (pinhole) in every test: the right-image grid is never consulted
    if (!grid) const_cast<Frame*>(this)->grid = ref_grid_build(mvKeysUn, mnMinX, mnMinY, mnMaxX, mnMaxY);
    return ref_grid_query(grid, mvKeysUn, x, y, r, minLevel, maxLevel);
  }
};

class KeyFrame {
 public:
  int N = 0, NLeft = -1, NRight = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
  std::vector<float> mvuRight, mvDepth;
  cv::Mat mDescriptors;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
  std::vector<MapPoint*> mvpMapPoints;
  float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0, mThDepth = 0;
  int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;      // const int in include/KeyFrame.h:352-355
  int mnScaleLevels = 8;
  float mfScaleFactor = 1.2f, mfLogScaleFactor = 0;
  std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
  cv::Mat Tcw, Ow, mTrl, mTlr;
  long unsigned int mnId = 0;
  RefGrid* grid = nullptr;
  std::vector<std::pair<MapPoint*, int>> addedMapPoints;   // log of AddMapPoint calls

  ~KeyFrame() { if (grid) ref_grid_free(grid); }
  cv::Mat GetPose() { return Tcw.clone(); }
  cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
  cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
  cv::Mat GetCameraCenter() { return Ow.clone(); }
  cv::Mat GetRightPose() { return mTrl * Tcw; }
  cv::Mat GetRightRotation() { return mTrl.rowRange(0, 3).colRange(0, 3) * Tcw.rowRange(0, 3).colRange(0, 3); }
  cv::Mat GetRightTranslation() { return mTrl.rowRange(0, 3).colRange(0, 3) * Tcw.rowRange(0, 3).col(3) + mTrl.rowRange(0, 3).col(3); }
  cv::Mat GetRightCameraCenter() { return Ow.clone(); }
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  std::set<MapPoint*> GetMapPoints() {
    std::set<MapPoint*> s;
    for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p);
    return s;
  }
  MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
  void AddMapPoint(MapPoint* pMP, const size_t& idx) { addedMapPoints.push_back(std::make_pair(pMP, (int)idx)); }
  bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }   // KeyFrame.cc:702-705
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const {
    (void)bRight;
    if (!grid) const_cast<KeyFrame*>(this)->grid = ref_grid_build(mvKeysUn, (float)mnMinX, (float)mnMinY, (float)mnMaxX, (float)mnMaxY);
    return ref_grid_query(grid, mvKeysUn, x, y, r, -1, -1);
  }
};

// src/MapPoint.cc:578-612
inline int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {
  float ratio = mfMaxDistance / currentDist;
  int nScale = ceil(log(ratio) / pKF->mfLogScaleFactor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
  return nScale;
}
inline int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
  float ratio = mfMaxDistance / currentDist;
  int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
  return nScale;
}

}  // namespace ORB_SLAM3
#endif
