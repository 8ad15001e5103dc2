// frame_prelude.h — TEST INFRASTRUCTURE.  Force-included (-include) in front of the reference's UNMODIFIED src/Frame.cc so
// that it compiles into oracle/_ref/libref_frame.so without Eigen / Boost / g2o / the rest of ORB-SLAM3.
//
// include/Frame.h itself is the reference's (class Frame is the real one); what it and Frame.cc pull in besides OpenCV is
// replaced here by pre-defining the include guards and supplying stand-ins with exactly the members Frame.cc touches:
//   ImuTypes.h        IMU::Bias / Calib / Preintegrated: state only (Calib's copy clones its matrices, src/ImuTypes.cc)
//   MapPoint.h        the tracking scratch isInFrustum writes + GetWorldPos / GetNormal / distance invariance / PredictScale
//   KeyFrame.h, G2oTypes.h (ConstraintPoseImu): pointers only
//   Converter.h       toDescriptorVector (src/Converter.cc:27-35), used by ComputeBoW only
//   ORBmatcher.h      TH_LOW / TH_HIGH (src/ORBmatcher.cc:36-37) and DescriptorDistance, which is forwarded to the
//                     reference's own compiled function in libref_matcher.so (resolved by the glue at load time)
//   CameraModels/*    Pinhole project / toK (src/CameraModels/Pinhole.cpp:31-41,148-153); KannalaBrandt8 aborts (out of scope)
// Every decision of ComputeStereoMatches, AssignFeaturesToGrid / GetFeaturesInArea, isInFrustum, UndistortKeyPoints,
// UpdatePoseMatrices, GetImu* and SetImuPoseVelocity is the reference's own compiled code.
#ifndef ORK_FRAME_PRELUDE_H_
#define ORK_FRAME_PRELUDE_H_
#define MAPPOINT_H
#define KEYFRAME_H
#define IMUTYPES_H
#define G2OTYPES_H
#define CONVERTER_H
#define ORBMATCHER_H
#define CAMERAMODELS_GEOMETRICCAMERA_H
#define CAMERAMODELS_PINHOLE_H
#define CAMERAMODELS_KANNALABRANDT8_H
#include "cvstub.h"
#include <climits>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

using namespace std;   // the reference headers replaced here leak it (include/Frame.h relies on that: `vector`, `map`, `string`)

namespace ORB_SLAM3 {

class Frame;
class KeyFrame {};
class ConstraintPoseImu {};

namespace IMU {
class Bias {
 public:
  float bax = 0, bay = 0, baz = 0, bwx = 0, bwy = 0, bwz = 0;
  Bias() {}
  Bias(float b_acc_x, float b_acc_y, float b_acc_z, float b_ang_vel_x, float b_ang_vel_y, float b_ang_vel_z)
      : bax(b_acc_x), bay(b_acc_y), baz(b_acc_z), bwx(b_ang_vel_x), bwy(b_ang_vel_y), bwz(b_ang_vel_z) {}
};
class Calib {   // include/ImuTypes.h:96-131
 public:
  cv::Mat Tcb, Tbc, Cov, CovWalk;
  Calib() {}
  Calib(const Calib& c) : Tcb(c.Tcb.clone()), Tbc(c.Tbc.clone()), Cov(c.Cov.clone()), CovWalk(c.CovWalk.clone()) {}
  Calib& operator=(const Calib& c) { Tcb = c.Tcb.clone(); Tbc = c.Tbc.clone(); Cov = c.Cov.clone(); CovWalk = c.CovWalk.clone(); return *this; }
};
class Preintegrated {
 public:
  Bias b;
  void SetNewBias(const Bias& bu) { b = bu; }
};
}  // namespace IMU

class GeometricCamera {   // Pinhole (src/CameraModels/Pinhole.cpp); arithmetic in float like the reference (mvParameters is vector<float>)
 public:
  float fx = 0, fy = 0, cx = 0, cy = 0;
  GeometricCamera() {}
  GeometricCamera(float fx_, float fy_, float cx_, float cy_) : fx(fx_), fy(fy_), cx(cx_), cy(cy_) {}
  virtual ~GeometricCamera() {}
  virtual cv::Point2f project(const cv::Point3f& p3D) { return cv::Point2f(fx * p3D.x / p3D.z + cx, fy * p3D.y / p3D.z + cy); }   // :31-34
  virtual cv::Point2f project(const cv::Mat& m3D) {                                                                             // :36-41
    const float* p3D = m3D.ptr<float>();
    return project(cv::Point3f(p3D[0], p3D[1], p3D[2]));
  }
  virtual cv::Mat toK() { return (cv::Mat_<float>(3, 3) << fx, 0.f, cx, 0.f, fy, cy, 0.f, 0.f, 1.f); }                         // :148-153
};
class Pinhole : public GeometricCamera {
 public:
  using GeometricCamera::GeometricCamera;
};
class KannalaBrandt8 : public GeometricCamera {
 public:
  float TriangulateMatches(GeometricCamera*, const cv::KeyPoint&, const cv::KeyPoint&, const cv::Mat&, const cv::Mat&, const float,
                           const float, cv::Mat&) { std::abort(); }
};

class MapPoint {
 public:
  // tracking scratch written by Frame::isInFrustum (include/MapPoint.h:133-146)
  float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
  bool mbTrackInView = false, mbTrackInViewR = false;
  int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
  float mTrackViewCos = 0, mTrackViewCosR = 0;
  long unsigned int mnId = 0;
  cv::Mat mWorldPos, mNormalVector;
  float mfMinDistance = 0, mfMaxDistance = 0;
  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
  cv::Mat GetNormal() { return mNormalVector.clone(); }
  float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }   // src/MapPoint.cc:566-570
  float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }   // src/MapPoint.cc:572-576
  int PredictScale(const float& currentDist, Frame* pF);              // src/MapPoint.cc:596-612 (defined by the glue, after Frame.h)
};

class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors) {   // src/Converter.cc:27-35
    std::vector<cv::Mat> vDesc;
    vDesc.reserve(Descriptors.rows);
    for (int j = 0; j < Descriptors.rows; j++) vDesc.push_back(Descriptors.row(j));
    return vDesc;
  }
};

extern "C" typedef int (*ref_descriptor_distance_fn)(const unsigned char*, const unsigned char*);
extern ref_descriptor_distance_fn g_ref_descriptor_distance;   // ORBmatcher::DescriptorDistance of libref_matcher.so
class ORBmatcher {
 public:
  static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;     // src/ORBmatcher.cc:36-38
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return g_ref_descriptor_distance(a.ptr(), b.ptr()); }
};

}  // namespace ORB_SLAM3
#endif
