// cvstub.cpp — TEST INFRASTRUCTURE: the numerical side of the OpenCV stand-in (cvstub.h).  Every primitive delegates
// to the oracle's restatement in ork_primitives.cpp, which tests/test_oracle_primitives.py pins bit-exactly to Python
// cv2 4.13; tests/test_ref_stub.py pins the float matrix algebra below to cv2 as well.
#include "cvstub.h"
#include "../ork.h"

namespace cv {

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& kps, int n) {
  // cv::KeyPointsFilter::retainBest: nth_element by response, then keep everything tied with the n-th
  if (n >= 0 && kps.size() > (size_t)n) {
    if (n == 0) { kps.clear(); return; }
    std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(),
                     [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    const float amb = kps[n - 1].response;
    auto end = std::partition(kps.begin() + n, kps.end(), [amb](const KeyPoint& k) { return k.response >= amb; });
    kps.resize(end - kps.begin());
  }
}

void Mat::copyTo(OutputArray dst) const {
  dst.create(rows, cols, flags_type);
  Mat d = dst.getMat();
  if (d.data == data && d.step.v == step.v) return;
  for (int y = 0; y < rows; ++y) std::memmove(d.data + (size_t)y * d.step.v, data + (size_t)y * step.v, (size_t)cols * elemSize());
}

static double get_elem(const Mat& m, int y, int x) {
  switch (m.type()) {
    case CV_8U: return m.at<uchar>(y, x);
    case CV_16S: return m.at<short>(y, x);
    case CV_16U: return m.at<unsigned short>(y, x);
    case CV_32S: return m.at<int>(y, x);
    case CV_32F: return m.at<float>(y, x);
    default: return m.at<double>(y, x);
  }
}
static void set_elem(Mat& m, int y, int x, double v) {
  switch (m.type()) {
    case CV_8U: m.at<uchar>(y, x) = saturate_cast<uchar>(v); break;
    case CV_16S: m.at<short>(y, x) = (short)std::min(32767.0, std::max(-32768.0, std::nearbyint(v))); break;
    case CV_16U: m.at<unsigned short>(y, x) = (unsigned short)std::min(65535.0, std::max(0.0, std::nearbyint(v))); break;
    case CV_32S: m.at<int>(y, x) = cvRound(v); break;
    case CV_32F: m.at<float>(y, x) = (float)v; break;
    default: m.at<double>(y, x) = v;
  }
}

void Mat::convertTo(Mat& dst, int type) const {
  Mat out(rows, cols, type);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) set_elem(out, y, x, get_elem(*this, y, x));
  dst = out;
}

Mat Mat::ones(int r, int c, int type) {
  Mat m(r, c, type);
  for (int y = 0; y < r; ++y)
    for (int x = 0; x < c; ++x) set_elem(m, y, x, 1.0);
  return m;
}
Mat Mat::eye(int r, int c, int type) {
  Mat m = zeros(r, c, type);
  for (int i = 0; i < std::min(r, c); ++i) set_elem(m, i, i, 1.0);
  return m;
}

// ---------------------------------------------------------------------------------------------------------------
// image primitives
// ---------------------------------------------------------------------------------------------------------------
void resize(InputArray src_, OutputArray dst_, Size dsize, double, double, int interpolation) {
  assert(interpolation == INTER_LINEAR);
  Mat src = src_.getMat();
  assert(src.type() == CV_8U);
  dst_.create(dsize.height, dsize.width, CV_8U);   // keeps a fitting ROI (the pyramid level inside its bordered buffer)
  Mat dst = dst_.getMat();
  ork::resize_linear_u8(src.data, src.cols, src.rows, (int)src.step.v, dst.data, dst.cols, dst.rows, (int)dst.step.v);
}

static inline int border101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
  return i;
}

void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType) {
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  Mat src = src_.getMat();
  dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
  Mat dst = dst_.getMat();
  // Without BORDER_ISOLATED OpenCV would read real pixels around an ROI; the reference only passes a non-isolated
  // source for level 0, whose source is a whole image, so both cases reduce to reflection about the ROI's own edges.
  const int w = src.cols, h = src.rows;
  std::vector<uchar> tmp((size_t)w * h);
  for (int y = 0; y < h; ++y) std::memcpy(tmp.data() + (size_t)y * w, src.ptr(y), w);   // src may live inside dst
  for (int y = 0; y < dst.rows; ++y) {
    const uchar* s = tmp.data() + (size_t)border101(y - top, h) * w;
    uchar* d = dst.ptr(y);
    for (int x = 0; x < dst.cols; ++x) d[x] = s[border101(x - left, w)];
  }
}

void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nms) {
  Mat img = image.getMat();
  assert(img.type() == CV_8U);
  std::vector<ork::FastPoint> pts;
  ork::fast9_16(img.data, img.cols, img.rows, (int)img.step.v, threshold, nms, pts);
  keypoints.clear();
  keypoints.reserve(pts.size());
  for (const ork::FastPoint& p : pts) keypoints.push_back(KeyPoint((float)p.x, (float)p.y, 7.f, -1.f, (float)p.score));
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType) {
  assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
  (void)ksize; (void)sigmaX; (void)sigmaY; (void)borderType;
  Mat src = src_.getMat();
  assert(src.type() == CV_8U);
  Mat out(src.rows, src.cols, CV_8U);
  ork::gaussian_blur7_s2(src.data, src.cols, src.rows, (int)src.step.v, out.data, (int)out.step.v);
  dst_.create(src.rows, src.cols, CV_8U);
  Mat dst = dst_.getMat();
  for (int y = 0; y < src.rows; ++y) std::memcpy(dst.ptr(y), out.ptr(y), src.cols);
}

float fastAtan2(float y, float x) { return ork::fast_atan2(y, x); }

// ---------------------------------------------------------------------------------------------------------------
// float / double matrix algebra, pinned to cv2 4.13 by tests/test_ref_stub.py.  cv::gemm has a small-matrix path:
// for CV_32F with inner dimension 2..4 equal to the result's width or height, every element is the plain fp32
// expression a0*b0 + a1*b1 (+ a2*b2 (+ a3*b3)) evaluated left to right (no FMA); everything else goes through
// GEMMSingleMul<float,double>: products and running sum in double, rounded to float once per element.
// ---------------------------------------------------------------------------------------------------------------
Mat operator*(const Mat& a, const Mat& b) {
  assert(a.cols == b.rows && a.type() == b.type() && (a.type() == CV_32F || a.type() == CV_64F));
  Mat c(a.rows, b.cols, a.type());
  const int len = a.cols;
  const bool small32 = a.type() == CV_32F && len >= 2 && len <= 4 && (len == c.cols || len == c.rows);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < b.cols; ++j) {
      if (small32) {
        volatile float s = a.at<float>(i, 0) * b.at<float>(0, j);   // volatile: one rounding per operation, whatever the flags
        for (int k = 1; k < len; ++k) { volatile float p = a.at<float>(i, k) * b.at<float>(k, j); s = s + p; }
        c.at<float>(i, j) = s;
      } else {
        double s = 0;
        for (int k = 0; k < len; ++k) s += get_elem(a, i, k) * get_elem(b, k, j);
        set_elem(c, i, j, s);
      }
    }
  return c;
}
template <typename F> static Mat elementwise(const Mat& a, const Mat& b, F f) {
  assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
  Mat c(a.rows, a.cols, a.type());
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols; ++j) {
      if (a.type() == CV_32F) c.at<float>(i, j) = f(a.at<float>(i, j), b.at<float>(i, j));
      else if (a.type() == CV_64F) c.at<double>(i, j) = f(a.at<double>(i, j), b.at<double>(i, j));
      else set_elem(c, i, j, f(get_elem(a, i, j), get_elem(b, i, j)));
    }
  return c;
}
Mat operator+(const Mat& a, const Mat& b) { return elementwise(a, b, [](auto x, auto y) { return x + y; }); }
Mat operator-(const Mat& a, const Mat& b) { return elementwise(a, b, [](auto x, auto y) { return x - y; }); }
Mat Mat::mul(const Mat& m) const { return elementwise(*this, m, [](auto x, auto y) { return x * y; }); }
Mat operator-(const Mat& a) { return a * -1.0; }
Mat operator*(const Mat& a, double s) {
  Mat c(a.rows, a.cols, a.type());
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols; ++j) {
      if (a.type() == CV_32F) c.at<float>(i, j) = (float)(a.at<float>(i, j) * s);   // cvt: saturate_cast<float>(src*alpha) in double
      else set_elem(c, i, j, get_elem(a, i, j) * s);
    }
  return c;
}
Mat operator*(double s, const Mat& a) { return a * s; }
Mat operator/(const Mat& a, double s) { return a * (1.0 / s); }

// Products with a lazily transposed operand: cv::gemm with GEMM_1_T / GEMM_2_T set skips the small-matrix path
// (its condition starts with flags == 0) and runs GEMMSingleMul<float,double>: products and running sum in double, k
// ascending, d = T(s * alpha).  Checked against cv2.gemm in tests/test_ref_stub.py.
static Mat gemm_general(const Mat& a, bool ta, const Mat& b, bool tb, double alpha) {
  const int ar = ta ? a.cols : a.rows, ac = ta ? a.rows : a.cols, br = tb ? b.cols : b.rows, bc = tb ? b.rows : b.cols;
  assert(ac == br && a.type() == b.type() && (a.type() == CV_32F || a.type() == CV_64F));
  (void)br;
  Mat c(ar, bc, a.type());
  for (int i = 0; i < ar; ++i)
    for (int j = 0; j < bc; ++j) {
      double s = 0;
      for (int k = 0; k < ac; ++k) s += (ta ? get_elem(a, k, i) : get_elem(a, i, k)) * (tb ? get_elem(b, j, k) : get_elem(b, k, j));
      set_elem(c, i, j, s * alpha);
    }
  return c;
}
MatT Mat::t() const { return MatT{*this, 1.0}; }
MatT::operator Mat() const { return alpha == 1.0 ? m.transposed() : m.transposed() * alpha; }
MatT operator-(const MatT& a) { return MatT{a.m, -a.alpha}; }
MatT operator*(double s, const MatT& a) { return MatT{a.m, a.alpha * s}; }
MatT operator*(const MatT& a, double s) { return MatT{a.m, a.alpha * s}; }
Mat operator*(const MatT& a, const Mat& b) { return gemm_general(a.m, true, b, false, a.alpha); }
Mat operator*(const Mat& a, const MatT& b) { return gemm_general(a, false, b.m, true, b.alpha); }
Mat operator*(const MatT& a, const MatT& b) { return gemm_general(a.m, true, b.m, true, a.alpha * b.alpha); }

Mat Mat::transposed() const {
  Mat c(cols, rows, flags_type);
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) std::memcpy(c.data + (size_t)j * c.step.v + (size_t)i * elemSize(), data + (size_t)i * step.v + (size_t)j * elemSize(), elemSize());
  return c;
}

double Mat::dot(const Mat& m) const {
  // dotProd_32f: double accumulator over float products taken in double
  double s = 0;
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) s += get_elem(*this, i, j) * get_elem(m, i, j);
  return s;
}

double norm(InputArray a_) {
  Mat a = a_.getMat();
  double s = 0;
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols; ++j) { const double v = get_elem(a, i, j); s += v * v; }
  return std::sqrt(s);
}
double norm(InputArray a, InputArray b) { return norm(a.getMat() - b.getMat()); }
double norm(InputArray a_, InputArray b_, int normType) {
  if (normType == NORM_L2) return norm(a_, b_);
  assert(normType == NORM_L1);
  const Mat a = a_.getMat(), b = b_.getMat();
  assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
  double s = 0;      // integer inputs: exact; cv::norm accumulates 16S differences in int and returns the sum as double
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols; ++j) s += std::fabs(get_elem(a, i, j) - get_elem(b, i, j));
  return s;
}
Mat operator-(const Mat& a, double sc) {
  Mat c(a.rows, a.cols, a.type());
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols; ++j) {
      double v = get_elem(a, i, j) - sc;
      if (a.type() == CV_16S) v = std::min(32767.0, std::max(-32768.0, std::nearbyint(v)));
      else if (a.type() == CV_8U) v = std::min(255.0, std::max(0.0, std::nearbyint(v)));
      set_elem(c, i, j, v);
    }
  return c;
}
extern "C" int ork_cv_undistort_points(const float* xy, int n, const orbx_camera* cam, const float* dist, int ndist, float* out_xy);
void undistortPoints(InputArray src_, OutputArray dst_, InputArray K_, InputArray D_, InputArray R_, InputArray P_) {
  const Mat src = src_.getMat(), K = K_.getMat(), D = D_.getMat(), P = P_.getMat();
  assert(src.type() == CV_32F && src.cols == 2 && R_.getMat().empty());
  // the oracle's restatement of cv::undistortPoints (K -> iterate -> P), pinned to cv2 4.13 by tests/test_oracle_frame.py;
  // the reference passes P = mK, the same intrinsics
  orbx_camera cam{};
  cam.fx = K.at<float>(0, 0); cam.fy = K.at<float>(1, 1); cam.cx = K.at<float>(0, 2); cam.cy = K.at<float>(1, 2);
  assert(P.empty() || (P.at<float>(0, 0) == cam.fx && P.at<float>(1, 1) == cam.fy && P.at<float>(0, 2) == cam.cx && P.at<float>(1, 2) == cam.cy));
  float d[5] = {0, 0, 0, 0, 0};
  const int nd = (int)D.total();
  for (int i = 0; i < std::min(nd, 5); ++i) d[i] = D.at<float>(i);
  std::vector<float> in((size_t)2 * src.rows), out((size_t)2 * src.rows);
  for (int i = 0; i < src.rows; ++i) { in[2 * i] = src.at<float>(i, 0); in[2 * i + 1] = src.at<float>(i, 1); }
  ork_cv_undistort_points(in.data(), src.rows, &cam, d, nd < 5 ? 4 : 5, out.data());
  Mat dst(src.rows, 2, CV_32F);
  for (int i = 0; i < src.rows; ++i) { dst.at<float>(i, 0) = out[2 * i]; dst.at<float>(i, 1) = out[2 * i + 1]; }
  dst_.getMatRef() = dst;
}

double determinant(InputArray a_) {
  Mat a = a_.getMat();
  assert(a.rows == a.cols);
  if (a.rows == 2) return get_elem(a, 0, 0) * get_elem(a, 1, 1) - get_elem(a, 0, 1) * get_elem(a, 1, 0);
  assert(a.rows == 3);
  auto e = [&](int i, int j) { return get_elem(a, i, j); };
  return e(0, 0) * (e(1, 1) * e(2, 2) - e(1, 2) * e(2, 1)) - e(0, 1) * (e(1, 0) * e(2, 2) - e(1, 2) * e(2, 0)) +
         e(0, 2) * (e(1, 0) * e(2, 1) - e(1, 1) * e(2, 0));
}

Mat Mat::inv(int) const {
  // cv::invert, DECOMP_LU: closed forms for 2x2 / 3x3 evaluated in double (Sf/Df macros of lapack.cpp), general case
  // Gauss-Jordan in double.
  assert(rows == cols);
  const int n = rows;
  Mat out(n, n, flags_type);
  auto e = [&](int i, int j) { return get_elem(*this, i, j); };
  if (n == 2) {
    double d = e(0, 0) * e(1, 1) - e(0, 1) * e(1, 0);
    if (d != 0) {
      d = 1. / d;
      const double t0 = e(0, 0) * d, t1 = e(1, 1) * d;
      set_elem(out, 1, 1, t0); set_elem(out, 0, 0, t1);
      set_elem(out, 0, 1, -e(0, 1) * d); set_elem(out, 1, 0, -e(1, 0) * d);
    }
    return out;
  }
  if (n == 3) {
    double d = determinant(*this);
    if (d != 0) {
      d = 1. / d;
      double t[9];
      t[0] = (e(1, 1) * e(2, 2) - e(1, 2) * e(2, 1)) * d;
      t[1] = (e(0, 2) * e(2, 1) - e(0, 1) * e(2, 2)) * d;
      t[2] = (e(0, 1) * e(1, 2) - e(0, 2) * e(1, 1)) * d;
      t[3] = (e(1, 2) * e(2, 0) - e(1, 0) * e(2, 2)) * d;
      t[4] = (e(0, 0) * e(2, 2) - e(0, 2) * e(2, 0)) * d;
      t[5] = (e(0, 2) * e(1, 0) - e(0, 0) * e(1, 2)) * d;
      t[6] = (e(1, 0) * e(2, 1) - e(1, 1) * e(2, 0)) * d;
      t[7] = (e(0, 1) * e(2, 0) - e(0, 0) * e(2, 1)) * d;
      t[8] = (e(0, 0) * e(1, 1) - e(0, 1) * e(1, 0)) * d;
      for (int i = 0; i < 9; ++i) set_elem(out, i / 3, i % 3, t[i]);
    }
    return out;
  }
  std::vector<double> A((size_t)n * 2 * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) A[(size_t)i * 2 * n + j] = e(i, j);
    A[(size_t)i * 2 * n + n + i] = 1;
  }
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r) if (std::fabs(A[(size_t)r * 2 * n + c]) > std::fabs(A[(size_t)p * 2 * n + c])) p = r;
    if (A[(size_t)p * 2 * n + c] == 0) return Mat::zeros(n, n, flags_type);
    if (p != c) for (int j = 0; j < 2 * n; ++j) std::swap(A[(size_t)p * 2 * n + j], A[(size_t)c * 2 * n + j]);
    const double d = 1. / A[(size_t)c * 2 * n + c];
    for (int j = 0; j < 2 * n; ++j) A[(size_t)c * 2 * n + j] *= d;
    for (int r = 0; r < n; ++r) if (r != c) {
      const double f = A[(size_t)r * 2 * n + c];
      if (f != 0) for (int j = 0; j < 2 * n; ++j) A[(size_t)r * 2 * n + j] -= f * A[(size_t)c * 2 * n + j];
    }
  }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) set_elem(out, i, j, A[(size_t)i * 2 * n + n + j]);
  return out;
}

void hconcat(InputArray a_, InputArray b_, OutputArray dst) {
  Mat a = a_.getMat(), b = b_.getMat();
  Mat out(a.rows, a.cols + b.cols, a.type());
  a.copyTo(out.colRange(0, a.cols));
  b.copyTo(out.colRange(a.cols, a.cols + b.cols));
  dst.getMatRef() = out;
}
void vconcat(InputArray a_, InputArray b_, OutputArray dst) {
  Mat a = a_.getMat(), b = b_.getMat();
  Mat out(a.rows + b.rows, a.cols, a.type());
  a.copyTo(out.rowRange(0, a.rows));
  b.copyTo(out.rowRange(a.rows, a.rows + b.rows));
  dst.getMatRef() = out;
}

std::ostream& operator<<(std::ostream& os, const Mat& m) {
  os << "[";
  for (int i = 0; i < m.rows; ++i) {
    for (int j = 0; j < m.cols; ++j) os << get_elem(m, i, j) << (j + 1 < m.cols ? ", " : "");
    os << (i + 1 < m.rows ? ";\n " : "");
  }
  return os << "]";
}

}  // namespace cv
