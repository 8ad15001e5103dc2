// ref_frame_glue.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's UNMODIFIED ORB_SLAM3::Frame
// (src/Frame.cc compiled in place behind frame_prelude.h) and ORB_SLAM3::ORBextractor (src/ORBextractor.cc), so that tests
// can build a real stereo / monocular Frame from Python and compare its ComputeStereoMatches, keypoint grid,
// isInFrustum, UndistortKeyPoints and IMU pose helpers with the oracle restatements.
#include "Frame.h"          // the reference's own header (-I /root/reference/include), behind frame_prelude.h
#include "ORBextractor.h"
#include <atomic>
#include <cstdint>
#include <cstring>
#include <dlfcn.h>
#include <new>

namespace ORB_SLAM3 {
ref_descriptor_distance_fn g_ref_descriptor_distance = nullptr;
// src/MapPoint.cc:596-612
int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
  float ratio = mfMaxDistance / currentDist;
  int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
  if (nScale < 0) nScale = 0;
  else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
  return nScale;
}
}  // namespace ORB_SLAM3

// Monotonic, thread-safe operator new while a Frame is being built: DistributeOctTree breaks ties by heap address
// (src/ORBextractor.cc:682); with addresses growing in allocation order the unmodified extractor follows the rule
// "later-created node = larger address" the oracle and the device implement (see ref_extractor_glue.cpp).  The two
// extraction threads of the stereo constructor (src/Frame.cc:111-114) share the arena; each thread's own allocations
// still come in increasing order.
namespace {
char* g_base = nullptr;
const size_t g_cap = (size_t)1 << 30;
std::atomic<size_t> g_used{0};
std::atomic<bool> g_active{false};
}
void* operator new(size_t n) {
  if (g_active.load(std::memory_order_relaxed)) {
    const size_t a = (n + 15) & ~(size_t)15;
    const size_t off = g_used.fetch_add(a, std::memory_order_relaxed);
    if (off + a <= g_cap) return g_base + off;
  }
  void* p = std::malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept {
  if (g_base && (char*)p >= g_base && (char*)p < g_base + g_cap) return;
  std::free(p);
}
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

using namespace ORB_SLAM3;

struct RefKeyPoint { float x, y, size, angle, response; int octave; };

struct RefFrame {
  ORBextractor *exL = nullptr, *exR = nullptr;
  Pinhole* cam = nullptr;
  Frame* F = nullptr;
  std::vector<MapPoint*> mps;
};

static cv::Mat wrap_u8(const uint8_t* img, int w, int h, int stride) { return cv::Mat(h, w, CV_8U, (void*)img, (size_t)stride); }

extern "C" {

// libref_matcher.so (the reference's compiled ORBmatcher.cc) supplies DescriptorDistance
int ref_frame_init(const char* matcher_lib) {
  if (ORB_SLAM3::g_ref_descriptor_distance) return 0;
  void* h = dlopen(matcher_lib, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1;
  ORB_SLAM3::g_ref_descriptor_distance = (ref_descriptor_distance_fn)dlsym(h, "ref_descriptor_distance");
  return ORB_SLAM3::g_ref_descriptor_distance ? 0 : -2;
}

// Frame::Frame(stereo) (src/Frame.cc:90-192) when imgR != NULL, else Frame::Frame(mono) (:308-384).  dist: 4 floats.
// The statics of class Frame (image bounds, grid cell sizes, fx ...) are (re)computed for this image size.
RefFrame* ref_frame_create(const uint8_t* imgL, const uint8_t* imgR, int w, int h, int stride, int nfeatures, float scaleFactor,
                           int nlevels, int iniTh, int minTh, float fx, float fy, float cx, float cy, const float* dist, float bf,
                           float thDepth, const float* Tcb /* 16 or NULL */) {
  if (!g_base) g_base = (char*)std::malloc(g_cap);
  RefFrame* r = new RefFrame();
  g_used = 0;
  g_active = true;
  r->exL = new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
  r->exR = imgR ? new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh) : nullptr;
  r->cam = new Pinhole(fx, fy, cx, cy);
  cv::Mat K = (cv::Mat_<float>(3, 3) << fx, 0.f, cx, 0.f, fy, cy, 0.f, 0.f, 1.f);
  cv::Mat D(4, 1, CV_32F);
  for (int i = 0; i < 4; ++i) D.at<float>(i) = dist ? dist[i] : 0.f;
  IMU::Calib calib;
  if (Tcb) {
    calib.Tcb = cv::Mat(4, 4, CV_32F);
    std::memcpy(calib.Tcb.data, Tcb, 64);
    calib.Tbc = calib.Tcb.inv();
  }
  Frame::mbInitialComputations = true;
  const cv::Mat L = wrap_u8(imgL, w, h, stride);
  if (imgR) {
    const cv::Mat R = wrap_u8(imgR, w, h, stride);
    r->F = new Frame(L, R, 0.0, r->exL, r->exR, nullptr, K, D, bf, thDepth, r->cam, nullptr, calib);
  } else {
    r->F = new Frame(L, 0.0, r->exL, nullptr, r->cam, D, bf, thDepth, nullptr, calib);
  }
  g_active = false;
  return r;
}
void ref_frame_destroy(RefFrame* r) {
  if (!r) return;
  for (MapPoint* p : r->mps) delete p;
  delete r->F; delete r->exL; delete r->exR; delete r->cam;
  delete r;
}
int ref_frame_n(RefFrame* r) { return r->F->N; }
int ref_frame_n_right(RefFrame* r) { return (int)r->F->mvKeysRight.size(); }
float ref_frame_mb(RefFrame* r) { return r->F->mb; }
float ref_frame_log_scale_factor(RefFrame* r) { return r->F->mfLogScaleFactor; }
void ref_frame_bounds(RefFrame* r, float* b4) { b4[0] = Frame::mnMinX; b4[1] = Frame::mnMinY; b4[2] = Frame::mnMaxX; b4[3] = Frame::mnMaxY; }

static void put_keys(const std::vector<cv::KeyPoint>& v, RefKeyPoint* out) {
  for (size_t i = 0; i < v.size(); ++i) out[i] = {v[i].pt.x, v[i].pt.y, v[i].size, v[i].angle, v[i].response, v[i].octave};
}
// which: 0 mvKeys, 1 mvKeysRight, 2 mvKeysUn
void ref_frame_keys(RefFrame* r, int which, RefKeyPoint* out, uint8_t* desc) {
  const Frame& F = *r->F;
  put_keys(which == 0 ? F.mvKeys : which == 1 ? F.mvKeysRight : F.mvKeysUn, out);
  if (desc) {
    const cv::Mat& d = which == 1 ? F.mDescriptorsRight : F.mDescriptors;
    for (int i = 0; i < d.rows; ++i) std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
  }
}

// Frame::ComputeStereoMatches (src/Frame.cc:955-1133).  The stereo constructor calls it BEFORE it assigns mb (:132 vs :166:
// minZ = mb is read uninitialised there); here it runs again on the finished Frame, where mb = mbf / fx holds.
void ref_frame_stereo_matches(RefFrame* r, float* uRight, float* depth) {
  Frame& F = *r->F;
  F.ComputeStereoMatches();
  for (int i = 0; i < F.N; ++i) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

// Frame::GetFeaturesInArea (src/Frame.cc:755-850) over the grid AssignFeaturesToGrid built (:444-478)
int ref_frame_features_in_area(RefFrame* r, float x, float y, float rad, int minLevel, int maxLevel, int32_t* out, int cap) {
  const std::vector<size_t> v = r->F->GetFeaturesInArea(x, y, rad, minLevel, maxLevel, false);
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int32_t)v[i];
  return (int)v.size();
}

// Frame::SetPose -> UpdatePoseMatrices (src/Frame.cc:489-544); outputs mOw, GetImuRotation, GetImuPosition (rig given at create)
void ref_frame_set_pose(RefFrame* r, const float* Tcw16, float* Ow3, float* Rwb9, float* twb3) {
  cv::Mat T(4, 4, CV_32F);
  std::memcpy(T.data, Tcw16, 64);
  r->F->SetPose(T);
  const cv::Mat Ow = r->F->GetCameraCenter();
  for (int i = 0; i < 3; ++i) Ow3[i] = Ow.at<float>(i);
  if (Rwb9 && !r->F->mImuCalib.Tcb.empty()) {
    const cv::Mat R = r->F->GetImuRotation(), t = r->F->GetImuPosition();
    for (int i = 0; i < 9; ++i) Rwb9[i] = R.at<float>(i / 3, i % 3);
    for (int i = 0; i < 3; ++i) twb3[i] = t.at<float>(i);
  }
}
// Frame::SetImuPoseVelocity (src/Frame.cc:520-530): inputs as Converter::toCvMat leaves them (float), output mTcw
void ref_frame_set_imu_pose(RefFrame* r, const float* Rwb9, const float* twb3, float* Tcw16) {
  cv::Mat R(3, 3, CV_32F), t(3, 1, CV_32F), v = cv::Mat::zeros(3, 1, CV_32F);
  std::memcpy(R.data, Rwb9, 36);
  for (int i = 0; i < 3; ++i) t.at<float>(i) = twb3[i];
  r->F->SetImuPoseVelocity(R, t, v);
  for (int i = 0; i < 16; ++i) Tcw16[i] = r->F->mTcw.at<float>(i / 4, i % 4);
}

// Frame::isInFrustum (src/Frame.cc:571-662, Nleft == -1) over n MapPoints at the pose last set
int ref_frame_is_in_frustum(RefFrame* r, int n, const float* xw, const float* maxDist, const float* minDist, const float* normal,
                            float viewingCosLimit, uint8_t* inView, float* projX, float* projY, float* projXR, float* depth,
                            int32_t* level, float* viewCos) {
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    MapPoint* p = new MapPoint();
    r->mps.push_back(p);
    p->mWorldPos = (cv::Mat_<float>(3, 1) << xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
    p->mNormalVector = (cv::Mat_<float>(3, 1) << normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
    p->mfMaxDistance = maxDist[i];              // raw mfMaxDistance / mfMinDistance (the getters apply 1.2 / 0.8)
    p->mfMinDistance = minDist[i];
    const bool in = r->F->isInFrustum(p, viewingCosLimit);
    inView[i] = in ? 1 : 0;
    projX[i] = p->mTrackProjX; projY[i] = p->mTrackProjY; projXR[i] = p->mTrackProjXR; depth[i] = p->mTrackDepth;
    level[i] = p->mnTrackScaleLevel; viewCos[i] = p->mTrackViewCos;
    cnt += in;
  }
  return cnt;
}

}  // extern "C"
