// ork.h — CPU ORACLE for the orbx hot path.  TEST INFRASTRUCTURE ONLY.
//
// This directory restates, on the CPU and without OpenCV/Eigen/g2o, the arithmetic of the
// reference's tracking hot path (file:line citations on every function).  It exists so the
// CUDA path can be checked bit-for-bit (integer/byte work) or to tolerance (fp64 optimisers).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it.  The product library (liborbx.so) never links, includes or calls anything here.
//
// PINNING.  The reference has no tests or golden vectors for this path and its CMake build cannot run
// here (OpenCV 3 / Eigen / Boost / Pangolin are not installed).  Since round 2 the oracle is pinned to REFERENCE SOURCE
// instead: oracle/_ref/ holds src/ORBextractor.cc, src/ORBmatcher.cc and Thirdparty/DBoW2 compiled UNMODIFIED against
// the OpenCV stand-in of oracle/ref_stub/, whose numerical primitives are pinned bit-exactly to Python cv2 4.13
// (tests/test_oracle_primitives.py, tests/test_ref_stub.py).  tests/test_oracle_vs_ref.py and tests/test_ref_matcher.py
// require oracle == reference source, bit for bit, on every input the GPU parity tests use (rows a1-a8, a10-a13, f1,
// f2).  Still PARITY UNPINNED: the optimisers (a15-a17, f3: g2o needs Eigen, which is not in the image), ComputeStereoMatches
// and Frame::GetFeaturesInArea / isInFrustum / UndistortKeyPoints (src/Frame.cc needs the whole of ORB-SLAM3's headers).
#ifndef ORK_H_
#define ORK_H_
#include <cstdint>
#include <cmath>
#include <vector>
#include "../include/orbx.h"  // orbx_keypoint POD only (shared wire struct, no code)

namespace ork {

// cvRound: round-half-to-even (OpenCV uses cvtss2si / lrint).
static inline int cv_round(float v) { return (int)std::lrintf(v); }
static inline int cv_round(double v) { return (int)std::lrint(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

struct Gray {  // un-bordered 8-bit image, tightly packed
  int w = 0, h = 0;
  std::vector<uint8_t> px;
  Gray() {}
  Gray(int w_, int h_) : w(w_), h(h_), px((size_t)w_ * h_) {}
  const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
  uint8_t* row(int y) { return px.data() + (size_t)y * w; }
};

struct FastPoint { int x, y, score; };

// ---- OpenCV primitives restated (cv2 4.13 semantics; see each .cpp comment) ----
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                      int dstride);
// cv::FAST(roi, th, nonmaxSuppression, TYPE_9_16); output raster order.
void fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, bool nms,
              std::vector<FastPoint>& out);
void gaussian_blur7_s2(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
float fast_atan2(float y, float x);

// ---- extractor (src/ORBextractor.cc) ----
struct Extractor {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;  // the reference stores the ctor's float in a double member (ORBextractor.h:97)
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> featuresPerLevel;
  std::vector<int> umax;
  std::vector<Gray> pyramid;                      // mvImagePyramid (without the unused border)
  std::vector<std::vector<orbx_keypoint>> cand;   // vToDistributeKeys per level (debug view)
  Extractor(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh);
  // returns 0, -1 (empty) or -2 (image too small / unsupported aspect)
  int extract(const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
              std::vector<orbx_keypoint>& kps, std::vector<uint8_t>& desc, int* monoIndex);
  void level_size(int w, int h, int level, int* lw, int* lh) const;
};

std::vector<orbx_keypoint> distribute_octree(const std::vector<orbx_keypoint>& keys, int minX, int maxX,
                                             int minY, int maxY, int N);
float ic_angle(const Gray& img, int x, int y, const std::vector<int>& umax);
void orb_descriptor(const Gray& blurred, int x, int y, float angleDeg, uint8_t* desc32);
extern const int8_t kPattern[1024];

}  // namespace ork
#endif
