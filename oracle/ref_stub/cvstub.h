// cvstub.h — TEST INFRASTRUCTURE: a minimal stand-in for the slice of the OpenCV C++ API that the reference's hot-path
// translation units use, so that those files compile UNMODIFIED from /root/reference into oracle/_ref/ (see
// oracle/Makefile, target _ref).  OpenCV's C++ headers and libraries are not installed in this image.
//
// What is restated here is interface only (cv::Mat as a ref-counted 2-D view, Point/Size/Rect/KeyPoint PODs,
// InputArray/OutputArray proxies).  The numerical primitives (resize, FAST, GaussianBlur, fastAtan2, copyMakeBorder,
// cvRound, float matrix algebra) are implemented in cvstub.cpp on top of the oracle's primitives, which are pinned
// bit-exactly to Python cv2 4.13 (tests/test_oracle_primitives.py, tests/test_ref_stub.py).
//
// Nothing under awesome-orb-slam3-3dvisioncraft-version_b200/ may include this file.
#ifndef ORK_CVSTUB_H_
#define ORK_CVSTUB_H_
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>
#include <iostream>
#include <limits>
#include <cstdlib>
#include <sstream>   // OpenCV's core.hpp pulls these in transitively; reference files rely on it
#include <fstream>
#include <map>
#include <list>
#include <set>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F
#define CV_32SC1 CV_32S
#define CV_PI 3.1415926535897932384626433832795
#define CV_INLINE static inline
#define CV_Assert(expr) assert(expr)

// cvRound: OpenCV uses cvtss2si / cvtsd2si (or lrint): round half to even.  Both overloads exist since 3.0.
CV_INLINE int cvRound(double v) { return (int)std::lrint(v); }
CV_INLINE int cvRound(float v) { return (int)std::lrintf(v); }
CV_INLINE int cvRound(int v) { return v; }
CV_INLINE int cvFloor(double v) { int i = (int)v; return i - (i > v); }
CV_INLINE int cvFloor(float v) { int i = (int)v; return i - (i > v); }
CV_INLINE int cvFloor(int v) { return v; }
CV_INLINE int cvCeil(double v) { int i = (int)v; return i + (i < v); }
CV_INLINE int cvCeil(float v) { int i = (int)v; return i + (i < v); }
CV_INLINE int cvCeil(int v) { return v; }

namespace cv {

using std::vector;
typedef std::string String;

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1 };
enum { NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& p) : x(saturate_cast<T>(p.x)), y(saturate_cast<T>(p.y)) {}
};
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b) {
  a.x = saturate_cast<T>(a.x * b);
  a.y = saturate_cast<T>(a.y * b);
  return a;
}
template <> inline Point_<float>& operator*=(Point_<float>& a, float b) { a.x = a.x * b; a.y = a.y * b; return a; }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;

template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Range {
  int start, end;
  Range() : start(0), end(0) {}
  Range(int s, int e) : start(s), end(e) {}
  static Range all() { return Range(INT32_MIN, INT32_MAX); }
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

struct MatStep {
  size_t v;
  MatStep() : v(0) {}
  MatStep(size_t s) : v(s) {}
  operator size_t() const { return v; }
};

class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
class MatExpr;

inline size_t cvstub_elem_size(int type) {
  switch (type) {
    case CV_8U: case CV_8S: return 1;
    case CV_16U: case CV_16S: return 2;
    case CV_32S: case CV_32F: return 4;
    default: return 8;
  }
}

// Single-channel 2-D matrix header over a ref-counted buffer: copying shares the data, ROI operators return views.
struct MatT;
class Mat {
 public:
  int flags_type, rows, cols;
  uchar* data;
  MatStep step;
  std::shared_ptr<std::vector<uchar>> buf;

  Mat() : flags_type(CV_8U), rows(0), cols(0), data(nullptr), step(0) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* ext, size_t st = 0) : flags_type(type), rows(r), cols(c), data((uchar*)ext),
        step(st ? st : c * cvstub_elem_size(type)) {}
  Mat(const Mat& m, const Rect& roi) : flags_type(m.flags_type), rows(roi.height), cols(roi.width),
        data(m.data + (size_t)roi.y * m.step.v + (size_t)roi.x * m.elemSize()), step(m.step), buf(m.buf) {}
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == flags_type) return;   // OpenCV keeps a fitting buffer
    flags_type = type; rows = r; cols = c;
    step = MatStep((size_t)c * cvstub_elem_size(type));
    buf = std::make_shared<std::vector<uchar>>((size_t)r * step.v + 64);
    data = buf->data();
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { buf.reset(); data = nullptr; rows = cols = 0; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return flags_type; }
  int depth() const { return flags_type; }
  int channels() const { return 1; }
  size_t elemSize() const { return cvstub_elem_size(flags_type); }
  size_t elemSize1() const { return elemSize(); }
  size_t step1() const { return step.v / elemSize(); }
  size_t total() const { return (size_t)rows * cols; }
  Size size() const { return Size(cols, rows); }
  bool isContinuous() const { return step.v == (size_t)cols * elemSize(); }
  Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
  Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
  Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
  Mat colRange(const Range& r) const { return colRange(r.start, r.end); }
  Mat row(int i) const { return rowRange(i, i + 1); }
  Mat col(int i) const { return colRange(i, i + 1); }
  Mat clone() const {
    Mat m(rows, cols, flags_type);
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step.v, data + (size_t)y * step.v, (size_t)cols * elemSize());
    return m;
  }
  void copyTo(OutputArray dst) const;
  void convertTo(Mat& dst, int type) const;
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step.v + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step.v + (size_t)x * sizeof(T)); }
  template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step.v; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step.v; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.v); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.v); }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); for (int y = 0; y < r; ++y) std::memset(m.ptr(y), 0, (size_t)c * m.elemSize()); return m; }
  static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
  static Mat ones(int r, int c, int type);
  static Mat eye(int r, int c, int type);
  // float / double algebra (see cvstub.cpp for the evaluation order, pinned to cv2 by tests/test_ref_stub.py).  t() is
  // the one lazy expression the stand-in keeps: cv::MatExpr turns A.t()*B and A*B.t() into cv::gemm calls with a
  // GEMM_x_T flag, and those never take gemm's small-matrix fp32 path.
  MatT t() const;
  Mat transposed() const;
  Mat reshape(int /*cn*/, int /*rows*/ = 0) const { return *this; }   // single-channel stand-in: N x 2 floats stay N x 2
  Mat inv(int method = DECOMP_LU) const;
  double dot(const Mat& m) const;
  Mat mul(const Mat& m) const;
  Mat& operator=(const Mat& m) = default;
  Mat(const Mat& m) = default;
};

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() : Mat() {}
  Mat_(int r, int c);
  Mat_(const Mat& m) : Mat(m) {}
  T& operator()(int y, int x) { return this->template at<T>(y, x); }
  const T& operator()(int y, int x) const { return this->template at<T>(y, x); }
};
template <> inline Mat_<float>::Mat_(int r, int c) : Mat(r, c, CV_32F) {}
template <> inline Mat_<double>::Mat_(int r, int c) : Mat(r, c, CV_64F) {}
template <> inline Mat_<uchar>::Mat_(int r, int c) : Mat(r, c, CV_8U) {}
template <> inline Mat_<int>::Mat_(int r, int c) : Mat(r, c, CV_32S) {}

// cv::Mat_<float>(3,1) << a, b, c   (MatCommaInitializer_)
template <typename T> class MatCommaInit {
 public:
  Mat_<T> m;
  size_t idx;
  MatCommaInit(const Mat_<T>& m_) : m(m_), idx(0) {}
  MatCommaInit& operator,(T v) { m.template at<T>((int)(idx / m.cols), (int)(idx % m.cols)) = v; ++idx; return *this; }
  operator Mat() const { return m; }
  operator Mat_<T>() const { return m; }
};
template <typename T, typename V> static inline MatCommaInit<T> operator<<(const Mat_<T>& m, V v) {
  MatCommaInit<T> ci(m);
  return (ci, (T)v);
}

class _InputArray {
 public:
  const Mat* m;
  Mat own;
  _InputArray() : m(nullptr) {}
  _InputArray(const Mat& mm) : m(&mm) {}
  template <typename T> _InputArray(const MatCommaInit<T>& ci) : m(nullptr), own(ci.m) { m = &own; }
  Mat getMat() const { return m ? *m : Mat(); }
  bool empty() const { return !m || m->empty(); }
};
class _OutputArray {
 public:
  Mat* m;
  _OutputArray() : m(nullptr) {}
  _OutputArray(Mat& mm) : m(&mm) {}
  _OutputArray(const Mat& mm) : m(const_cast<Mat*>(&mm)) {}
  Mat getMat() const { return m ? *m : Mat(); }
  Mat& getMatRef() const { return *m; }
  void create(int r, int c, int type) const { if (m) m->create(r, c, type); }
  void create(Size s, int type) const { if (m) m->create(s, type); }
  void release() const { if (m) m->release(); }
  bool needed() const { return m != nullptr; }
};
inline InputArray noArray() { static _InputArray none; return none; }

// cv::FileStorage / cv::FileNode: DBoW2's TemplatedVocabulary has virtual save/load members over them, so they must
// compile; the hot path only uses loadFromTextFile / loadFromBinaryFile.  Calling any of these aborts.
class FileNode {
 public:
  FileNode operator[](const char*) const { std::abort(); }
  FileNode operator[](const std::string&) const { std::abort(); }
  FileNode operator[](int) const { std::abort(); }
  size_t size() const { std::abort(); }
  operator int() const { std::abort(); }
  operator float() const { std::abort(); }
  operator double() const { std::abort(); }
  operator std::string() const { std::abort(); }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) { std::abort(); }
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const char*) const { std::abort(); }
  FileNode operator[](const std::string&) const { std::abort(); }
};
template <typename T> static inline FileStorage& operator<<(FileStorage& fs, const T&) { std::abort(); return fs; }

// ---- the primitives (cvstub.cpp) ----
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
float fastAtan2(float y, float x);
double norm(InputArray a);
double norm(InputArray a, InputArray b);
double norm(InputArray a, InputArray b, int normType);   // NORM_L1 / NORM_L2 of a - b
// N x 2 CV_32F points (the reference reshapes to 2 channels and back around the call: reshape() is the identity here)
void undistortPoints(InputArray src, OutputArray dst, InputArray K, InputArray distCoeffs, InputArray R, InputArray P);
struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0; };
class BFMatcher {   // Frame::BFmatcher: only ComputeStereoFishEyeMatches (KannalaBrandt8 rigs, out of scope) calls it
 public:
  BFMatcher(int = NORM_L2, bool = false) {}
  void knnMatch(InputArray, InputArray, std::vector<std::vector<DMatch>>&, int) const { std::abort(); }
};
double determinant(InputArray a);
void hconcat(InputArray a, InputArray b, OutputArray dst);
void vconcat(InputArray a, InputArray b, OutputArray dst);

// A.t() (optionally scaled: -A.t(), s*A.t()) as an operand; converts to the materialised (scaled) transpose anywhere else
struct MatT {
  Mat m;
  double alpha;
  operator Mat() const;
  Mat inv(int method = DECOMP_LU) const { return Mat(*this).inv(method); }
};
MatT operator-(const MatT& a);
MatT operator*(double s, const MatT& a);
MatT operator*(const MatT& a, double s);
Mat operator*(const MatT& a, const Mat& b);     // gemm(..., GEMM_1_T)
Mat operator*(const Mat& a, const MatT& b);     // gemm(..., GEMM_2_T)
Mat operator*(const MatT& a, const MatT& b);    // gemm(..., GEMM_1_T | GEMM_2_T)
Mat operator*(const Mat& a, const Mat& b);
Mat operator+(const Mat& a, const Mat& b);
Mat operator-(const Mat& a, const Mat& b);
Mat operator-(const Mat& a);
Mat operator-(const Mat& a, double s);           // Mat - Scalar, saturating for the integer types
Mat operator*(const Mat& a, double s);
Mat operator*(double s, const Mat& a);
Mat operator/(const Mat& a, double s);
std::ostream& operator<<(std::ostream& os, const Mat& m);

}  // namespace cv
#endif
