// ref_iface.h — SYNTAX-CHECK STAND-INS, not product code and not the reference's headers.
//
// Declares just enough of OpenCV's and the reference's public interface (names, member types, signatures as
// used by the shim) for `g++ -fsyntax-only` to type-check the shim sources in a container that has neither
// OpenCV nor the reference's dependencies.  Interface facts restated from include/ORBextractor.h:43-109,
// include/ORBmatcher.h:35-108, include/Optimizer.h:46-119, include/Frame.h, include/KeyFrame.h,
// include/MapPoint.h, include/Map.h of the reference.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {
struct Point2f { float x, y; };
struct Rect { int x, y, w, h; Rect(int a, int b, int c, int d) : x(a), y(b), w(c), h(d) {} };
struct KeyPoint {
  Point2f pt; float size, angle, response; int octave, class_id;
  KeyPoint() {}
  KeyPoint(float x, float y, float s, float a, float r, int o, int c) : pt{x, y}, size(s), angle(a), response(r), octave(o), class_id(c) {}
};
struct Mat {
  int rows, cols; size_t step; unsigned char* data;
  Mat() {}
  Mat(int r, int c, int type);
  Mat operator()(const Rect&) const;
  Mat rowRange(int, int) const; Mat colRange(int, int) const; Mat inv(int method = 0) const;
  int type() const;
  bool empty() const;
  template <typename T> T& at(int r, int c = 0);
  template <typename T> const T& at(int r, int c = 0) const;
  template <typename T = unsigned char> T* ptr(int r = 0);
  template <typename T = unsigned char> const T* ptr(int r = 0) const;
};
struct InputArray { bool empty() const; Mat getMat() const; };
struct OutputArray { void release() const; void create(int, int, int) const; Mat getMat() const; };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { DECOMP_SVD = 1 };
void copyMakeBorder(const Mat&, Mat&, int, int, int, int, int);
}  // namespace cv

namespace DBoW2 {
typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;
typedef std::map<unsigned int, double> BowVector;
}

namespace Eigen {
template <typename T, int R, int C> struct Matrix {
  T& operator()(int r, int c); const T& operator()(int r, int c) const;
  T& operator()(int i); const T& operator()(int i) const;
};
typedef Matrix<double, 3, 3> Matrix3d; typedef Matrix<double, 3, 1> Vector3d; typedef Matrix<double, 2, 1> Vector2d;
}

namespace ORB_SLAM3 {
class Map; class KeyFrame; class Frame;
typedef Eigen::Matrix<double, 9, 9> Matrix9d; typedef Eigen::Matrix<double, 15, 15> Matrix15d;
class GeometricCamera { public: float uncertainty2(const Eigen::Vector2d& p2D); };
namespace IMU {                                          // include/ImuTypes.h
struct Bias { Bias() {} Bias(float ax, float ay, float az, float wx, float wy, float wz); float bax, bay, baz, bwx, bwy, bwz; };
struct Calib { cv::Mat Tcb, Tbc; };
class Preintegrated {
 public:
  cv::Mat GetDeltaRotation(const Bias& b); cv::Mat GetDeltaVelocity(const Bias& b); cv::Mat GetDeltaPosition(const Bias& b);
  Bias GetOriginalBias(); float dT; cv::Mat C, dR, dV, dP, JRg, JVg, JVa, JPg, JPa;
};
}  // namespace IMU
class ImuCamPose {                                       // include/G2oTypes.h:58-103
 public:
  Eigen::Vector3d twb; Eigen::Matrix3d Rwb;          // (per-camera Rcw/tcw/Rcb/tcb/Rbc/tbc vectors are not read by the shim)
};
class VertexPose { public: VertexPose(Frame*); VertexPose(KeyFrame*); const ImuCamPose& estimate() const; };   // :106-137
class EdgeInertial { public: EdgeInertial(IMU::Preintegrated*); const Matrix9d& information() const; };         // :497-545
class ConstraintPoseImu {                                // :704-749
 public:
  ConstraintPoseImu(const Eigen::Matrix3d& Rwb, const Eigen::Vector3d& twb, const Eigen::Vector3d& vwb, const Eigen::Vector3d& bg,
                    const Eigen::Vector3d& ba, const Matrix15d& H);
  Eigen::Matrix3d Rwb; Eigen::Vector3d twb, vwb, bg, ba; Matrix15d H;
};
namespace Converter { cv::Mat toCvMat(const Eigen::Matrix3d&); cv::Mat toCvMat(const Eigen::Vector3d&); }
class ORBVocabulary;   // DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> in the reference (include/ORBVocabulary.h:36)

class ORBextractor {
 public:
  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~ORBextractor() {}
  int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                 cv::OutputArray _descriptors, std::vector<int>& vLappingArea);
  std::vector<cv::Mat> mvImagePyramid;
 protected:
  int nfeatures; double scaleFactor; int nlevels; int iniThFAST; int minThFAST;
  std::vector<int> mnFeaturesPerLevel, umax;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

class MapPoint {
 public:
  cv::Mat GetWorldPos(); void SetWorldPos(const cv::Mat&); cv::Mat GetDescriptor();
  std::map<KeyFrame*, std::tuple<int, int> > GetObservations(); int Observations();
  void EraseObservation(KeyFrame*); bool isBad(); void UpdateNormalAndDepth(); Map* GetMap();
  bool IsInKeyFrame(KeyFrame*); cv::Mat GetNormal(); void Replace(MapPoint*); void AddObservation(KeyFrame*, int);
  float GetMaxDistance(); float GetMinDistance();   // one-line accessors of mfMaxDistance / mfMinDistance
  long unsigned int mnId, mnBALocalForKF;
  float mTrackProjX, mTrackProjY, mTrackDepth, mTrackProjXR, mTrackViewCos; bool mbTrackInView; int mnTrackScaleLevel;
  static std::mutex mGlobalMutex;
};

class Frame {
 public:
  void SetPose(cv::Mat Tcw); void ComputeStereoMatches(); void ComputeBoW(); void UndistortKeyPoints();
  int isInFrustumBatch(const std::vector<MapPoint*>& vpMP, float viewingCosLimit);   // added member
  void SetImuPoseVelocity(const cv::Mat& Rwb, const cv::Mat& twb, const cv::Mat& Vwb);
  KeyFrame* mpLastKeyFrame; Frame* mpPrevFrame; IMU::Preintegrated *mpImuPreintegrated, *mpImuPreintegratedFrame; IMU::Calib mImuCalib; IMU::Bias mImuBias; cv::Mat mVw;
  ConstraintPoseImu* mpcpi; GeometricCamera* mpCamera;
  cv::Mat mDistCoef, mRcw, mtcw, mOw; int mnScaleLevels; float mfLogScaleFactor;
  ORBVocabulary* mpORBvocabulary; DBoW2::BowVector mBowVec; DBoW2::FeatureVector mFeatVec; int Nleft;
  ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
  static float fx, fy, cx, cy; float mbf, mb; int N;
  std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn; std::vector<MapPoint*> mvpMapPoints;
  std::vector<float> mvuRight, mvDepth; cv::Mat mDescriptors, mDescriptorsRight; std::vector<bool> mvbOutlier;
  cv::Mat mTcw; std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};

class KeyFrame {
 public:
  cv::Mat GetPose(); cv::Mat GetRotation(); cv::Mat GetTranslation(); void SetPose(const cv::Mat&);
  std::vector<KeyFrame*> GetVectorCovisibleKeyFrames(); std::vector<MapPoint*> GetMapPointMatches();
  MapPoint* GetMapPoint(const size_t& idx); void EraseMapPointMatch(MapPoint*); bool isBad(); Map* GetMap();
  cv::Mat GetCameraCenter(); void AddMapPoint(MapPoint*, const size_t& idx); void ComputeBoW();
  ORBVocabulary* mpORBvocabulary; DBoW2::BowVector mBowVec; float mfLogScaleFactor; int N, NLeft;
  cv::Mat GetVelocity(); IMU::Bias GetImuBias();
  long unsigned int mnId, mnBALocalForKF, mnBAFixedForKF;
  const float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
  std::vector<cv::KeyPoint> mvKeysUn; std::vector<float> mvuRight; cv::Mat mDescriptors; DBoW2::FeatureVector mFeatVec;
  std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  const int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
};

class Map { public: long unsigned int GetInitKFid(); bool IsInertial(); std::mutex mMutexMapUpdate; };

class ORBmatcher {
 public:
  ORBmatcher(float nnratio = 0.6, bool checkOri = true);
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false, const float thFarPoints = 50.0f);
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
  int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
  int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0, const bool bRight = false);
  int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo, const bool bCoarse = false);
 protected:
  float mfNNratio; bool mbCheckOrientation;
};

class Optimizer {
 public:
  static void LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, int& num_fixedKF);
  static int PoseOptimization(Frame* pFrame);
  static int PoseInertialOptimizationLastKeyFrame(Frame* pFrame, bool bRecInit = false);
  static int PoseInertialOptimizationLastFrame(Frame* pFrame, bool bRecInit = false);
};
}  // namespace ORB_SLAM3
