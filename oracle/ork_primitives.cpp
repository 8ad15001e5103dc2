// ork_primitives.cpp — ORACLE (test infrastructure).  The four OpenCV primitives the reference's
// extractor delegates to, restated as plain integer / fp32 code and pinned bit-exactly to
// Python cv2 4.13 by tests/test_oracle_primitives.py.
//
//   cv::resize(INTER_LINEAR, 8U)   called at  src/ORBextractor.cc:1171
//   cv::FAST(roi, th, true)        called at  src/ORBextractor.cc:808,827
//   cv::GaussianBlur(7x7, s=2)     called at  src/ORBextractor.cc:1121
//   cv::fastAtan2                  called at  src/ORBextractor.cc:101
//
// OpenCV itself is not vendored in the reference (CMakeLists.txt:36-43 asks for OpenCV 3, README
// says 3.2); the published fixed-point algorithms of OpenCV 4.x are restated here because cv2
// 4.13 is the only executable OpenCV available to pin against.
#include "ork.h"
#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace ork {

// ---------------------------------------------------------------------------------------------
// Bilinear resize, 8-bit, 11-bit fixed-point coefficients (INTER_RESIZE_COEF_BITS = 11).
// For every destination index d along an axis: f = (float)((d+0.5)*scale-0.5), s=floor(f),
// f-=s.  Horizontally OpenCV forces (s<0 -> s=0,f=0) and (s>=src-1 -> s=src-1,f=0);
// vertically it only clips the two row indices.  Coefficients are saturate_cast<short>(w*2048)
// (round-half-even).  Horizontal pass is exact int32; the vertical pass is the 8-bit
// specialisation  ((b0*(T0>>4))>>16) + ((b1*(T1>>4))>>16) + 2) >> 2.
// ---------------------------------------------------------------------------------------------
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                      int dstride) {
  const double inv_sx = (double)dw / sw, inv_sy = (double)dh / sh;
  const double scale_x = 1.0 / inv_sx, scale_y = 1.0 / inv_sy;
  std::vector<int> xofs(dw), yofs0(dh), yofs1(dh);
  std::vector<short> xa0(dw), xa1(dw), yb0(dh), yb1(dh);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cv_floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    xa0[dx] = (short)cv_round((1.f - fx) * 2048.f);
    xa1[dx] = (short)cv_round(fx * 2048.f);
  }
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor(fy);
    fy -= sy;
    yofs0[dy] = std::min(std::max(sy, 0), sh - 1);
    yofs1[dy] = std::min(std::max(sy + 1, 0), sh - 1);
    yb0[dy] = (short)cv_round((1.f - fy) * 2048.f);
    yb1[dy] = (short)cv_round(fy * 2048.f);
  }
  std::vector<int> T0(dw), T1(dw);
  for (int dy = 0; dy < dh; ++dy) {
    const uint8_t* S0 = src + (size_t)yofs0[dy] * sstride;
    const uint8_t* S1 = src + (size_t)yofs1[dy] * sstride;
    for (int dx = 0; dx < dw; ++dx) {
      int x0 = xofs[dx], x1 = std::min(x0 + 1, sw - 1);
      T0[dx] = S0[x0] * xa0[dx] + S0[x1] * xa1[dx];
      T1[dx] = S1[x0] * xa0[dx] + S1[x1] * xa1[dx];
    }
    uint8_t* D = dst + (size_t)dy * dstride;
    const int b0 = yb0[dy], b1 = yb1[dy];
    for (int dx = 0; dx < dw; ++dx)
      D[dx] = (uint8_t)((((b0 * (T0[dx] >> 4)) >> 16) + ((b1 * (T1[dx] >> 4)) >> 16) + 2) >> 2);
  }
}

// ---------------------------------------------------------------------------------------------
// FAST-9-16.  Ring (dx,dy), clockwise from 6 o'clock as OpenCV enumerates it.  With
// d_k = I(p) - I(ring_k): arcmax = max over the 16 cyclic arcs of 9 consecutive ring pixels of
// max(min d, min -d).  p is a corner iff arcmax > threshold; its NMS score is arcmax-1 (what
// cornerScore<16> returns for a corner, independent of the threshold).  Evaluated only on
// [3,w-3)x[3,h-3) of the ROI; NMS keeps p iff score(p) > score(q) for its 8 neighbours, where
// non-corner and un-evaluated neighbours count 0.  Output in raster order.
// ---------------------------------------------------------------------------------------------
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// Returns arcmax when it exceeds t; otherwise some value <= t (early-out on the opposite-pair test:
// every 9-arc contains one pixel of each antipodal pair, so a corner needs, for one polarity, a
// pixel beyond the threshold in all 8 pairs).
static inline int fast_arcmax(const uint8_t* p, const int* off, int t) {
  const int v = p[0];
  int d[25];
  d[0] = v - p[off[0]];
  d[8] = v - p[off[8]];
  bool br = (d[0] > t) | (d[8] > t), dk = (d[0] < -t) | (d[8] < -t);
  if (!(br | dk)) return t;
  d[4] = v - p[off[4]];
  d[12] = v - p[off[12]];
  br &= (d[4] > t) | (d[12] > t);
  dk &= (d[4] < -t) | (d[12] < -t);
  if (!(br | dk)) return t;
  for (int k = 1; k < 8; ++k) {
    if (k == 4) continue;
    d[k] = v - p[off[k]];
    d[k + 8] = v - p[off[k + 8]];
    br &= (d[k] > t) | (d[k + 8] > t);
    dk &= (d[k] < -t) | (d[k + 8] < -t);
  }
  if (!(br | dk)) return t;
  for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
  int best = -256;
  for (int k = 0; k < 16; ++k) {
    int mn = d[k], mx = d[k];
    for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

void fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, bool nms,
              std::vector<FastPoint>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  int off[16];
  for (int k = 0; k < 16; ++k) off[k] = kRingDy[k] * stride + kRingDx[k];
  // score map: -1 = not a corner (a corner at threshold 0 may legitimately score 0)
  std::vector<int16_t> score((size_t)w * h, (int16_t)-1);
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      int a = fast_arcmax(img + (size_t)y * stride + x, off, threshold);
      if (a > threshold) {
        score[(size_t)y * w + x] = (int16_t)(a - 1);
        if (!nms) out.push_back({x, y, 0});
      }
    }
  if (!nms) return;
  auto S = [&](const int16_t* s) { return *s < 0 ? 0 : (int)*s; };
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      const int16_t* s = &score[(size_t)y * w + x];
      if (*s < 0) continue;
      int sc = *s;
      if (sc > S(s - 1) && sc > S(s + 1) && sc > S(s - w - 1) && sc > S(s - w) && sc > S(s - w + 1) &&
          sc > S(s + w - 1) && sc > S(s + w) && sc > S(s + w + 1))
        out.push_back({x, y, sc});
    }
}

// ---------------------------------------------------------------------------------------------
// GaussianBlur 8U, 7x7, sigma 2, BORDER_REFLECT_101 as computed by OpenCV 4.x's fixed-point
// path: separable 8.8 kernel [18,34,48,56,48,34,18] (sum 256), no rounding between passes,
// dst = (sum_v k_v * (sum_u k_u I) + 2^15) >> 16.
// ---------------------------------------------------------------------------------------------
static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
  return i;
}

void gaussian_blur7_s2(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
  std::vector<uint32_t> hbuf((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* S = src + (size_t)y * sstride;
    for (int x = 0; x < w; ++x) {
      uint32_t acc = 0;
      for (int k = 0; k < 7; ++k) acc += K[k] * S[reflect101(x + k - 3, w)];
      hbuf[(size_t)y * w + x] = acc;
    }
  }
  for (int y = 0; y < h; ++y) {
    uint8_t* D = dst + (size_t)y * dstride;
    for (int x = 0; x < w; ++x) {
      uint32_t acc = 0;
      for (int k = 0; k < 7; ++k) acc += K[k] * hbuf[(size_t)reflect101(y + k - 3, h) * w + x];
      D[x] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cv::fastAtan2 (degrees): 7th-order odd polynomial on min/max ratio, plain fp32, no FMA
// (this TU is compiled with -ffp-contract=off).
// ---------------------------------------------------------------------------------------------
float fast_atan2(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + eps);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + eps);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

}  // namespace ork
